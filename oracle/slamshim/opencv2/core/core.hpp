/* oracle/slamshim: stand-in for <opencv2/core/core.hpp> */
#pragma once
#include "slamshim_cv.h"
