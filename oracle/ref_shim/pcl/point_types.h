// oracle/ref_shim: stands in for <pcl/point_types.h>, which /root/reference/include/kfusion/internal.hpp:6 includes
// without using anything from it in the device-side declarations.  Test infrastructure only.
#pragma once
