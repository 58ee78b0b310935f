#include "../../shim_matcher/cvm.hpp"
