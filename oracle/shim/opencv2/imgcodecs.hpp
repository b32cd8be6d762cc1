#include "../pvo_shim_cv.hpp"
