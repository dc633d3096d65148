// Empty stand-in: the reference's Utils.hpp:10 includes this header and uses nothing from it.
#pragma once
