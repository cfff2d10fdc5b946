// TEST INFRASTRUCTURE ONLY: forwards <tbb/parallel_for.h> to the std::thread stand-in (see tbb_shim.h).
#include "tbb_shim.h"
