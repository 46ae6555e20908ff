// TEST INFRASTRUCTURE: see path.hpp in this directory.
#include "path.hpp"
