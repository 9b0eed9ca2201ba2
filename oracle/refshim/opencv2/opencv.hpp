// stand-in (TEST INFRASTRUCTURE, see oracle/refshim/README.md)
#include "core.hpp"
