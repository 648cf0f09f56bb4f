// TEST INFRASTRUCTURE (oracle/): the two PCL types the reference's map_manager.hpp names (a point cloud kept for visualisation), so that
// the reference's headers and sources compile here without ROS / PCL.  Nothing on the checked paths reads the cloud.
#pragma once
#include <memory>
#include <vector>

namespace pcl {
struct PointXYZ { float x = 0, y = 0, z = 0; };
struct PointXYZRGB {
    float x = 0, y = 0, z = 0;
    unsigned char r = 0, g = 0, b = 0;
    PointXYZRGB() {}
    PointXYZRGB(unsigned char r_, unsigned char g_, unsigned char b_) : r(r_), g(g_), b(b_) {}
};
template <class P> struct PointCloud {
    typedef std::shared_ptr<PointCloud> Ptr;
    std::vector<P> points;
    void push_back(const P& p) { points.push_back(p); }
    size_t size() const { return points.size(); }
    void clear() { points.clear(); }
    void reserve(size_t n) { points.reserve(n); }
    bool empty() const { return points.empty(); }
};
}  // namespace pcl
