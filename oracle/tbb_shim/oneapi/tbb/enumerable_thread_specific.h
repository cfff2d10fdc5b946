// TEST INFRASTRUCTURE ONLY: forwards <oneapi/tbb/enumerable_thread_specific.h> to the std::thread stand-in.
#include "../../tbb/tbb_shim.h"
