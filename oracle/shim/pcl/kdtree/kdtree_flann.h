#include "../../pvo_shim_pcl.hpp"
