// TEST INFRASTRUCTURE ONLY: forwards <tbb/parallel_sort.h> to the std::thread stand-in (see tbb_shim.h).
#include "tbb_shim.h"
