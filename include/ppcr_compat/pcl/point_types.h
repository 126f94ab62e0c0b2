// Minimal pcl::PointXYZ: the 16-byte, 16-byte-aligned record of PCL (x, y, z + one float of padding).
#ifndef PPCR_COMPAT_PCL_POINT_TYPES_H
#define PPCR_COMPAT_PCL_POINT_TYPES_H
namespace pcl {
struct alignas(16) PointXYZ {
    union {
        float data[4];
        struct {
            float x, y, z;
        };
    };
    PointXYZ() : data{0.f, 0.f, 0.f, 1.f} {}
    PointXYZ(float x_, float y_, float z_) : data{x_, y_, z_, 1.f} {}
};
static_assert(sizeof(PointXYZ) == 16, "pcl::PointXYZ is a 16-byte record");
}  // namespace pcl
#endif
