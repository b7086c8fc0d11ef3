/* TEST INFRASTRUCTURE ONLY. Force-included when compiling the reference's
 * models/EMD/cuda/emd_kernel.cu against torch >= 2.x: glog's CHECK_EQ is no longer
 * provided by the torch headers, so map it onto TORCH_CHECK. Nothing else is changed. */
#pragma once
#include <c10/util/Exception.h>
#ifndef CHECK_EQ
#define CHECK_EQ(a, b) TORCH_CHECK((a) == (b), "CHECK_EQ failed: " #a " == " #b)
#endif
