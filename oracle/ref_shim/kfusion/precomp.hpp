// oracle/ref_shim: shadows /root/reference/include/kfusion/precomp.hpp for the reference's .cu files.
// The original pulls the whole host side (OpenCV, PCL, kinfu.hpp); its .cu translation units only need the CUDA
// vector helpers and kfusion/internal.hpp (the device-side declarations), which is what this keeps.
// Test infrastructure only: lets the reference's kernels compile *where they lie* into oracle/_ref/.
#pragma once
#include <vector_functions.h>
#include <climits>
#include <iostream>
#include <kfusion/internal.hpp>
