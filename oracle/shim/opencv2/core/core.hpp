// TEST INFRASTRUCTURE: see opencv2/imgproc/imgproc.hpp in this directory.
#include "../imgproc/imgproc.hpp"
