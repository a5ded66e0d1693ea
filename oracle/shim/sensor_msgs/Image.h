#pragma once
#include <memory>
#include <vector>
#include "../std_msgs/Header.h"
namespace sensor_msgs {
struct Image { std_msgs::Header header; unsigned height = 0, width = 0; std::string encoding; std::vector<unsigned char> data; typedef std::shared_ptr<Image> Ptr; typedef std::shared_ptr<const Image> ConstPtr; };
typedef std::shared_ptr<Image> ImagePtr;
namespace image_encodings { static const std::string BGR8 = "bgr8", MONO8 = "mono8"; }
}  // namespace sensor_msgs
