// include/Compare.h -- device-side comparison of two CSR matrices (reference include/Compare.h,
// source/GPU/Compare.cu:66-82): row lengths and positional column ids, optionally values.
#pragma once
#include "dCSR.h"

namespace spECK {
template <typename DataType>
bool Compare(const dCSR<DataType> &reference_mat, const dCSR<DataType> &compare_mat, bool compare_data);
}
