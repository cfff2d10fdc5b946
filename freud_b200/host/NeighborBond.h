// freud::locality::NeighborBond lives in NeighborQuery.h; the reference's binding layer includes "NeighborBond.h"
// (freud/locality/export-NeighborList.cc:11).
#pragma once
#include "NeighborQuery.h"
