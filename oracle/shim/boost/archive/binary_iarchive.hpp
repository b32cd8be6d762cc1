#include "binary_oarchive.hpp"
