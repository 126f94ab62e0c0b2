// OutputStream -- prints to std::cout only when verbose (reference: .../output_stream.hpp:7-23).
#ifndef PROB_POINT_CLOUD_REGISTRATION_OUTPUT_STREAM_HPP
#define PROB_POINT_CLOUD_REGISTRATION_OUTPUT_STREAM_HPP
#include <iostream>
namespace prob_point_cloud_registration {
class OutputStream {
public:
    explicit OutputStream(bool verbose = false) : verbose_(verbose) {}
    template <typename T>
    OutputStream& operator<<(const T& value)
    {
        if (verbose_) std::cout << value;
        return *this;
    }

private:
    bool verbose_;
};
}  // namespace prob_point_cloud_registration
#endif
