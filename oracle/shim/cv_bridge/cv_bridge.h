// ORACLE — TEST INFRASTRUCTURE ONLY.  cv_bridge names used by the reference's visualisation code; nothing is converted.
#pragma once
#include <opencv2/core/core.hpp>
#include "../sensor_msgs/Image.h"
namespace cv_bridge {
struct CvImage {
  std_msgs::Header header; std::string encoding; cv::Mat image;
  CvImage() {}
  CvImage(const std_msgs::Header& h, const std::string& e, const cv::Mat& m) : header(h), encoding(e), image(m) {}
  sensor_msgs::ImagePtr toImageMsg() const { return sensor_msgs::ImagePtr(new sensor_msgs::Image()); }
};
}  // namespace cv_bridge
