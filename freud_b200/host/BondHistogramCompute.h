// freud::locality::BondHistogramCompute lives in RDF.h (its first client); the reference's binding layer includes
// "BondHistogramCompute.h" (freud/locality/export-BondHistogramCompute.cc, freud/density/export-RDF.cc:9).
#pragma once
#include "RDF.h"
