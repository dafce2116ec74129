// TEST INFRASTRUCTURE ONLY: abseil stand-in, see oracle/absl_shim/shim_all.h
#include "shim_all.h"
