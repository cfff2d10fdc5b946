// TEST INFRASTRUCTURE ONLY: forwards <tbb/global_control.h> to the std::thread stand-in (see tbb_shim.h).
#include "tbb_shim.h"
