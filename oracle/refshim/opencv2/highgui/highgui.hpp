// stand-in (TEST INFRASTRUCTURE, see oracle/refshim/README.md)
#include "opencv2/core.hpp"
