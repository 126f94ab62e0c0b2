// Minimal pcl::PointCloud<T>: points vector, width/height/is_dense, shared_ptr typedefs (PCL >= 1.11).
#ifndef PPCR_COMPAT_PCL_POINT_CLOUD_H
#define PPCR_COMPAT_PCL_POINT_CLOUD_H
#include <cstddef>
#include <cstdint>
#include <memory>
#include <cstdlib>
#include <new>
#include <vector>

// Optional: page-locked storage from libppcr_cuda (include/ppcr.h).  Weak, so that this header also works in programs
// that do not link the CUDA library (the allocator then is plain malloc).
extern "C" {
void* ppcr_host_alloc(std::size_t bytes) __attribute__((weak));
std::int32_t ppcr_host_free(void* p) __attribute__((weak));
}

namespace pcl {
namespace detail {
// Large clouds (a PCD file read by loadPCDFile, the copy the registration makes) live in page-locked memory when the CUDA
// library is present: the host-to-device copy of ppcr_create then runs asynchronously at link speed instead of being staged
// through the driver's bounce buffers.  Small ones are not worth a cudaHostAlloc.
template <typename T>
struct HostAllocator {
    using value_type = T;
    static constexpr std::size_t kPinFromBytes = 1u << 20;
    HostAllocator() = default;
    template <typename U>
    HostAllocator(const HostAllocator<U>&) {}
    T* allocate(std::size_t n)
    {
        const std::size_t bytes = n * sizeof(T);
        if (bytes >= kPinFromBytes && ppcr_host_alloc)
            if (void* p = ppcr_host_alloc(bytes)) return static_cast<T*>(p);
        void* p = std::malloc(bytes ? bytes : 1);
        if (!p) throw std::bad_alloc();
        return static_cast<T*>(p);
    }
    void deallocate(T* p, std::size_t) noexcept
    {
        if (ppcr_host_free && ppcr_host_free(p)) return;
        std::free(p);
    }
    template <typename U>
    bool operator==(const HostAllocator<U>&) const { return true; }
    template <typename U>
    bool operator!=(const HostAllocator<U>&) const { return false; }
};
}  // namespace detail

template <typename PointT>
class PointCloud {
public:
    using Ptr = std::shared_ptr<PointCloud<PointT>>;
    using ConstPtr = std::shared_ptr<const PointCloud<PointT>>;
    using VectorType = std::vector<PointT, detail::HostAllocator<PointT>>;  // (PCL: std::vector<PointT, Eigen::aligned_allocator<PointT>>)
    using iterator = typename VectorType::iterator;
    using const_iterator = typename VectorType::const_iterator;
    VectorType points;
    std::uint32_t width = 0;
    std::uint32_t height = 0;
    bool is_dense = true;
    std::size_t size() const { return points.size(); }
    bool empty() const { return points.empty(); }
    void clear() { points.clear(); width = height = 0; }
    void resize(std::size_t n) { points.resize(n); width = static_cast<std::uint32_t>(n); height = 1; }
    void push_back(const PointT& p) { points.push_back(p); width = static_cast<std::uint32_t>(points.size()); height = 1; }
    PointT& operator[](std::size_t i) { return points[i]; }
    const PointT& operator[](std::size_t i) const { return points[i]; }
    PointT& at(std::size_t i) { return points.at(i); }
    const PointT& at(std::size_t i) const { return points.at(i); }
    iterator begin() { return points.begin(); }
    iterator end() { return points.end(); }
    const_iterator begin() const { return points.begin(); }
    const_iterator end() const { return points.end(); }
};
}  // namespace pcl
#endif
