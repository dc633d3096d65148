// Empty stand-in: test/test_spmv.cpp:8 includes <boost/filesystem.hpp> and uses nothing from it.
#pragma once
