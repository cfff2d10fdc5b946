// TEST INFRASTRUCTURE ONLY: forwards <tbb/concurrent_hash_map.h> to the std::thread stand-in (see tbb_shim.h).
#include "tbb_shim.h"
