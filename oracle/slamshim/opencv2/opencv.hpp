#pragma once
#include "../slamshim_cv.h"
