// Reference build of Optimizer.cc (oracle/_ref/libref_optimizer.so): the data-only stand-in for S/include/Frame.h is the mock the drop-in tests use.
// Test infrastructure only.
#pragma once
#include "../../orbslamm_b200/host/mock/Frame.h"
