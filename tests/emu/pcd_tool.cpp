// pcd_tool <in.pcd> <out.pcd> [binary] -- loads a PCD through the product's reader (include/ppcr_compat/pcl/io) and
// writes it back through its writer.  Test infrastructure for tests/test_cli.py.
#include <iostream>

#include <pcl/io/pcd_io.h>
#include <pcl/point_types.h>

int main(int argc, char** argv)
{
    if (argc < 3) return 2;
    pcl::PointCloud<pcl::PointXYZ> cloud;
    if (pcl::io::loadPCDFile<pcl::PointXYZ>(argv[1], cloud) == -1) {
        std::cout << "load failed" << std::endl;
        return 1;
    }
    const bool binary = argc > 3 && std::string(argv[3]) == "binary";
    if (pcl::io::savePCDFile(argv[2], cloud, binary) == -1) return 3;
    std::cout << cloud.size() << std::endl;
    return 0;
}
