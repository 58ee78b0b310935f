#include "cv_real_shim.hpp"
