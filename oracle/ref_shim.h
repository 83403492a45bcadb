// TEST INFRASTRUCTURE (oracle). Not part of the product.
// Force-included when compiling the UNMODIFIED reference sources (where they lie under
// /root/reference) for the CPU (-DHAS_NO_CUDA) build: the reference's own cudaStubs header
// lacks two symbols that tfQMRgpu/source/tfqmrgpu.cu:573,581 uses, and relies on transitive includes.
#pragma once
#include <cstdint>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cassert>
#include <vector>
#include <algorithm>
#define cudaFuncSetAttribute(...) ((void)0)
#define cudaFuncAttributeMaxDynamicSharedMemorySize 0
