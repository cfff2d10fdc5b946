// TEST INFRASTRUCTURE ONLY: forwards <oneapi/tbb/parallel_sort.h> to the std::thread stand-in.
#include "../../tbb/tbb_shim.h"
