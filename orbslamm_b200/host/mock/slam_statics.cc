// definitions of the mock data model's statics (test scaffolding)
#include "Frame.h"
namespace iORB_SLAM
{
std::mutex MapPoint::mGlobalMutex;
float Frame::fx, Frame::fy, Frame::cx, Frame::cy;
float Frame::mnMinX, Frame::mnMaxX, Frame::mnMinY, Frame::mnMaxY;
}
