/* oracle/cvshim: stand-in for <opencv2/core/core.hpp> (see ../cvshim.h or ../../cvshim.h) */
#pragma once
#include "cvshim.h"
