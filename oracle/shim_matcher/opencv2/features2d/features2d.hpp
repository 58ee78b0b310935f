#include "../../cvm.hpp"
