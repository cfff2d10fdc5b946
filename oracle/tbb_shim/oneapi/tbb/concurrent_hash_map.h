// TEST INFRASTRUCTURE ONLY: forwards <oneapi/tbb/concurrent_hash_map.h> to the std::thread stand-in.
#include "../../tbb/tbb_shim.h"
