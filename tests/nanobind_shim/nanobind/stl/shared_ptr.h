// TEST INFRASTRUCTURE ONLY -- see nanobind.h in this directory.
#pragma once
