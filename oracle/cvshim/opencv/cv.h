/* oracle/cvshim: stand-in for <opencv/cv.h> (see ../cvshim.h or ../../cvshim.h) */
#pragma once
#include "cvshim.h"
