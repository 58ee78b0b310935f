// TEST INFRASTRUCTURE: forwards to the from-scratch OpenCV-API shim (oracle/shim/cvshim.hpp).
#include "cvshim.hpp"
