// see opencv.hpp in this directory
#include "opencv.hpp"
