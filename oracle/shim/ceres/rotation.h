#include "../pvo_shim_ceres.hpp"
