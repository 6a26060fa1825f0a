// oracle/ntl_shim/NTL/LLL.h - TEST INFRASTRUCTURE ONLY (see RR.h).  common/PolyUpdate.cpp:3 includes it and uses
// nothing from it.
#pragma once
#include "ZZ.h"
#include "mat_RR.h"
