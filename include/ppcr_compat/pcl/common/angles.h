#include <pcl/common/transforms.h>
