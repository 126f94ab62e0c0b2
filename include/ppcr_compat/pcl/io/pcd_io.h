// Minimal PCD reader / writer for pcl::PointXYZ clouds: the on-disk format either side of the registration
// (src/prob_point_cloud_registration_ex.cc:113,123,132,164 in the reference).
//   loadPCDFile : DATA ascii | binary | binary_compressed; any field list containing x, y, z stored as 4-byte floats
//                 (other fields are skipped).  Returns 0, or -1 on any error like PCL.
//   savePCDFile : ASCII by default, like pcl::io::savePCDFile(name, cloud, binary_mode = false).
#ifndef PPCR_COMPAT_PCL_PCD_IO_H
#define PPCR_COMPAT_PCL_PCD_IO_H
#include <cmath>
#include <cstdint>
#include <exception>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <limits>
#include <sstream>
#include <string>
#include <vector>

#include <pcl/point_cloud.h>
#include <pcl/point_types.h>

namespace pcl {
namespace io {
namespace detail {

struct PcdField {
    std::string name;
    int size = 4;
    char type = 'F';
    int count = 1;
    int offset = 0;  // byte offset inside one point record
};

// LZF decompression (the codec of DATA binary_compressed)
inline bool lzf_decompress(const unsigned char* in, std::size_t in_len, unsigned char* out, std::size_t out_len)
{
    std::size_t ip = 0, op = 0;
    while (ip < in_len) {
        unsigned ctrl = in[ip++];
        if (ctrl < 32) {  // literal run
            const std::size_t n = ctrl + 1;
            if (op + n > out_len || ip + n > in_len) return false;
            std::memcpy(out + op, in + ip, n);
            op += n;
            ip += n;
        } else {  // back reference
            std::size_t len = ctrl >> 5;
            if (len == 7) {
                if (ip >= in_len) return false;
                len += in[ip++];
            }
            if (ip >= in_len) return false;
            const std::size_t dist = ((ctrl & 0x1f) << 8) + in[ip++] + 1;
            len += 2;
            if (dist > op || op + len > out_len) return false;
            for (std::size_t k = 0; k < len; ++k, ++op) out[op] = out[op - dist];
        }
    }
    return op == out_len;
}

inline float read_as_float(const unsigned char* p, const PcdField& f)
{
    if (f.type == 'F' && f.size == 4) {
        float v;
        std::memcpy(&v, p, 4);
        return v;
    }
    if (f.type == 'F' && f.size == 8) {
        double v;
        std::memcpy(&v, p, 8);
        return static_cast<float>(v);
    }
    return 0.f;
}

}  // namespace detail

template <typename PointT>
int loadPCDFile(const std::string& file_name, PointCloud<PointT>& cloud)
{
    std::ifstream f(file_name, std::ios::binary);
    if (!f) return -1;
    std::vector<detail::PcdField> fields;
    std::size_t width = 0, height = 1, n_points = 0;
    bool have_points = false;
    std::string data_kind, line;
    while (std::getline(f, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        if (line.empty() || line[0] == '#') continue;
        std::istringstream ls(line);
        std::string key;
        ls >> key;
        if (key == "VERSION") continue;
        if (key == "FIELDS" || key == "COLUMNS") {
            std::string name;
            while (ls >> name) {
                detail::PcdField fd;
                fd.name = name;
                fields.push_back(fd);
            }
        } else if (key == "SIZE") {
            for (auto& fd : fields)
                if (!(ls >> fd.size)) return -1;
        } else if (key == "TYPE") {
            for (auto& fd : fields)
                if (!(ls >> fd.type)) return -1;
        } else if (key == "COUNT") {
            for (auto& fd : fields)
                if (!(ls >> fd.count)) return -1;
        } else if (key == "WIDTH") {
            ls >> width;
        } else if (key == "HEIGHT") {
            ls >> height;
        } else if (key == "VIEWPOINT") {
            continue;
        } else if (key == "POINTS") {
            ls >> n_points;
            have_points = true;
        } else if (key == "DATA") {
            ls >> data_kind;
            break;
        }
    }
    if (fields.empty() || data_kind.empty()) return -1;
    if (!have_points) {
        if (height != 0 && width > std::numeric_limits<std::size_t>::max() / height) return -1;
        n_points = width * height;
    }
    // Header values come from the file: nothing below may index or allocate on their say-so alone.
    constexpr std::size_t kMaxRecord = std::size_t(1) << 20, kMaxPoints = std::size_t(1) << 31;
    if (n_points > kMaxPoints) return -1;
    int ix = -1, iy = -1, iz = -1, offset = 0;
    for (std::size_t k = 0; k < fields.size(); ++k) {
        if (fields[k].size != 1 && fields[k].size != 2 && fields[k].size != 4 && fields[k].size != 8) return -1;
        if (fields[k].count < 1 || static_cast<std::size_t>(fields[k].count) > kMaxRecord) return -1;
        if (static_cast<std::size_t>(offset) + static_cast<std::size_t>(fields[k].size) * fields[k].count > kMaxRecord) return -1;
        fields[k].offset = offset;
        offset += fields[k].size * fields[k].count;
        if (fields[k].name == "x") ix = static_cast<int>(k);
        if (fields[k].name == "y") iy = static_cast<int>(k);
        if (fields[k].name == "z") iz = static_cast<int>(k);
    }
    if (ix < 0 || iy < 0 || iz < 0) return -1;
    const std::size_t record = static_cast<std::size_t>(offset);
    if (data_kind != "ascii") {
        // a binary body holds record * n_points bytes (at most that, compressed): a POINTS value the file cannot back is rejected
        // before anything is allocated for it
        const std::streampos body = f.tellg();
        f.seekg(0, std::ios::end);
        const std::streampos end = f.tellg();
        f.seekg(body);
        if (body < 0 || end < body) return -1;
        const std::size_t left = static_cast<std::size_t>(end - body);
        if (data_kind == "binary" && (record == 0 || n_points > left / record)) return -1;
        if (data_kind == "binary_compressed" && left < 8) return -1;
    }
    try {
        cloud.points.assign(n_points, PointT());
    } catch (const std::exception&) {
        return -1;
    }
    cloud.width = static_cast<std::uint32_t>(width ? width : n_points);
    cloud.height = static_cast<std::uint32_t>(height);
    cloud.is_dense = true;
    // like PCL's reader: a cloud holding a non-finite coordinate is flagged, so that filters and the search skip those points
    auto flag_density = [&cloud]() {
        for (const auto& p : cloud.points)
            if (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z)) {
                cloud.is_dense = false;
                return;
            }
    };
    if (data_kind == "ascii") {
        for (std::size_t i = 0; i < n_points; ++i) {
            if (!std::getline(f, line)) return -1;
            std::istringstream ls(line);
            for (std::size_t k = 0; k < fields.size(); ++k)
                for (int c = 0; c < fields[k].count; ++c) {
                    std::string tok;
                    if (!(ls >> tok)) return -1;
                    if (c > 0) continue;
                    float v;
                    if (tok == "nan" || tok == "NaN" || tok == "-nan") {
                        v = std::numeric_limits<float>::quiet_NaN();
                        cloud.is_dense = false;
                    } else {
                        try {
                            v = std::stof(tok);
                        } catch (...) {
                            return -1;
                        }
                    }
                    if (static_cast<int>(k) == ix) cloud.points[i].x = v;
                    if (static_cast<int>(k) == iy) cloud.points[i].y = v;
                    if (static_cast<int>(k) == iz) cloud.points[i].z = v;
                }
        }
        flag_density();
        return 0;
    }
    std::vector<unsigned char> raw;
    try {
        raw.resize(record * n_points);
    } catch (const std::exception&) {
        return -1;
    }
    if (data_kind == "binary") {
        f.read(reinterpret_cast<char*>(raw.data()), static_cast<std::streamsize>(raw.size()));
        if (static_cast<std::size_t>(f.gcount()) != raw.size()) return -1;
        for (std::size_t i = 0; i < n_points; ++i) {
            const unsigned char* p = raw.data() + i * record;
            cloud.points[i].x = detail::read_as_float(p + fields[ix].offset, fields[ix]);
            cloud.points[i].y = detail::read_as_float(p + fields[iy].offset, fields[iy]);
            cloud.points[i].z = detail::read_as_float(p + fields[iz].offset, fields[iz]);
        }
        flag_density();
        return 0;
    }
    if (data_kind == "binary_compressed") {
        std::uint32_t comp = 0, uncomp = 0;
        f.read(reinterpret_cast<char*>(&comp), 4);
        f.read(reinterpret_cast<char*>(&uncomp), 4);
        if (!f || uncomp != raw.size()) return -1;
        {
            const std::streampos at = f.tellg();
            f.seekg(0, std::ios::end);
            const std::streampos end = f.tellg();
            f.seekg(at);
            if (at < 0 || end < at || static_cast<std::size_t>(end - at) < comp) return -1;
        }
        std::vector<unsigned char> packed(comp);
        f.read(reinterpret_cast<char*>(packed.data()), comp);
        if (static_cast<std::size_t>(f.gcount()) != comp) return -1;
        if (!detail::lzf_decompress(packed.data(), comp, raw.data(), raw.size())) return -1;
        // compressed files are stored field by field (structure of arrays)
        std::size_t base = 0;
        for (std::size_t k = 0; k < fields.size(); ++k) {
            const std::size_t fsz = static_cast<std::size_t>(fields[k].size) * fields[k].count;
            if (static_cast<int>(k) == ix || static_cast<int>(k) == iy || static_cast<int>(k) == iz) {
                for (std::size_t i = 0; i < n_points; ++i) {
                    const float v = detail::read_as_float(raw.data() + base + i * fsz, fields[k]);
                    if (static_cast<int>(k) == ix) cloud.points[i].x = v;
                    if (static_cast<int>(k) == iy) cloud.points[i].y = v;
                    if (static_cast<int>(k) == iz) cloud.points[i].z = v;
                }
            }
            base += fsz * n_points;
        }
        flag_density();
        return 0;
    }
    return -1;
}

template <typename PointT>
int savePCDFile(const std::string& file_name, const PointCloud<PointT>& cloud, bool binary_mode = false)
{
    std::ofstream f(file_name, std::ios::binary);
    if (!f) return -1;
    const std::size_t n = cloud.points.size();
    f << "# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\n"
      << "WIDTH " << n << "\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS " << n << "\nDATA " << (binary_mode ? "binary" : "ascii")
      << "\n";
    if (binary_mode) {
        for (const auto& p : cloud.points) {
            const float v[3] = {p.x, p.y, p.z};
            f.write(reinterpret_cast<const char*>(v), 12);
        }
    } else {
        f << std::setprecision(8);
        for (const auto& p : cloud.points) f << p.x << ' ' << p.y << ' ' << p.z << '\n';
    }
    return f ? 0 : -1;
}

template <typename PointT>
int savePCDFileBinary(const std::string& file_name, const PointCloud<PointT>& cloud)
{
    return savePCDFile(file_name, cloud, true);
}

}  // namespace io
}  // namespace pcl
#endif
