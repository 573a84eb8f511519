/* oracle/cvshim: stand-in for <opencv2/features2d/features2d.hpp> (see ../cvshim.h or ../../cvshim.h) */
#pragma once
#include "cvshim.h"
