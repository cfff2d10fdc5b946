// TEST INFRASTRUCTURE ONLY: forwards <oneapi/tbb/concurrent_vector.h> to the std::thread stand-in.
#include "../../tbb/tbb_shim.h"
