// TEST INFRASTRUCTURE ONLY: forwards <tbb/enumerable_thread_specific.h> to the std::thread stand-in (see tbb_shim.h).
#include "tbb_shim.h"
