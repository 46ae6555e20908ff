// TEST INFRASTRUCTURE: see opencv2/imgproc/imgproc.hpp in this directory.
#ifndef PHOVO_SHIM_OPENCV_CONTRIB_HPP_
#define PHOVO_SHIM_OPENCV_CONTRIB_HPP_
#include "../imgproc/imgproc.hpp"
namespace cv
{
class TickMeter
{
public:
  void start() {}
  void stop() {}
  double getTimeSec() const { return 0.; }
};
}
#endif
