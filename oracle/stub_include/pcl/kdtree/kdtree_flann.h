// Stand-in (TEST INFRASTRUCTURE): the reference's nanoflann wrapper includes this PCL header but uses nothing of it.
#ifndef DDLO_ORACLE_PCL_KDTREE_FLANN_STUB
#define DDLO_ORACLE_PCL_KDTREE_FLANN_STUB
#include <pcl/point_cloud.h>
#endif
