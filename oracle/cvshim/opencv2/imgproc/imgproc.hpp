/* oracle/cvshim: stand-in for <opencv2/imgproc/imgproc.hpp> (see ../cvshim.h or ../../cvshim.h) */
#pragma once
#include "cvshim.h"
