// Compiled as C++17 by tests/test_abi.py: the header-only C++ mirror builds against the C-ABI.
#include <cstdio>
#include "../wax_b200/host/cuda_vector_engine.hpp"
int main() {
    try {
        wax::CUDAVectorEngine bad(wax::VectorMetric::cosine, 0);
        return 1;
    } catch (const wax::InvalidToc &e) {
        std::printf("InvalidToc: %s available=%d\n", e.what(), wax::CUDAVectorEngine::isAvailable() ? 1 : 0);
    }
    return 0;
}
