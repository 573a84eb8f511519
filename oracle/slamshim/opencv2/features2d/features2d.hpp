/* oracle/slamshim: stand-in for <opencv2/features2d/features2d.hpp> */
#pragma once
#include "slamshim_cv.h"
