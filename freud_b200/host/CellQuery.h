// freud::locality::CellQuery lives in NeighborQuery.h with the other engines; this header exists so that the reference's
// binding layer, which includes "CellQuery.h" (freud/locality/export-NeighborQuery.cc:9-14), finds the replacement class and
// not the stock one.
#pragma once
#include "NeighborQuery.h"
