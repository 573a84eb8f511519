/* oracle/cvshim: stand-in for <opencv2/highgui/highgui.hpp> (see ../cvshim.h or ../../cvshim.h) */
#pragma once
#include "cvshim.h"
