// TEST INFRASTRUCTURE (oracle/): nothing of this OpenCV header is used on the compiled path (see core.hpp)
#pragma once
#include <opencv2/core.hpp>
#include <opencv2/imgproc.hpp>
#include <opencv2/features2d.hpp>
