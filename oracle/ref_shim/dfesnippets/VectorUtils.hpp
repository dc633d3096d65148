// Empty stand-in for the absent lib/dfe-snippets submodule (Spmv.cpp:5 includes, uses nothing).
#pragma once
