// GPU test of the header-only C++ mirror (wax_b200/host/cuda_vector_engine.hpp) over the C-ABI: the reference's
// own engine tests (Tests/WaxIntegrationTests/VectorSearchEngineTests.swift:7-47) written against the C++ surface.
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "../wax_b200/host/cuda_vector_engine.hpp"

#define EXPECT(cond)                                                            \
    do {                                                                        \
        if (!(cond)) { std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); return 1; } \
    } while (0)

using wax::CUDAVectorEngine;
using wax::VectorMetric;

static bool contains(const std::vector<CUDAVectorEngine::Hit> &hits, uint64_t id) {
    for (auto &h : hits) if (h.first == id) return true;
    return false;
}

int main() {
    if (!CUDAVectorEngine::isAvailable()) { std::printf("no CUDA device\n"); return 2; }
    {   // vectorEngineAddSearchRemoveRoundtrip (:7-19)
        CUDAVectorEngine engine(VectorMetric::cosine, 4);
        engine.add(0, {1, 0, 0, 0});
        engine.add(1, {0, 1, 0, 0});
        auto hits = engine.search({1, 0, 0, 0}, 10);
        EXPECT(!hits.empty() && contains(hits, 0) && hits[0].first == 0 && std::fabs(hits[0].second - 1.0f) < 1e-6f);
        engine.remove(0);
        EXPECT(!contains(engine.search({1, 0, 0, 0}, 10), 0));
    }
    {   // vectorEngineSerializeDeserializeRoundtripPreservesSearch (:21-34)
        CUDAVectorEngine engine(VectorMetric::cosine, 4);
        engine.add(0, {1, 0, 0, 0});
        engine.add(1, {0, 1, 0, 0});
        auto blob = engine.serialize();
        EXPECT(blob.size() == 36 + 2 * 16 + 8 + 16 && blob[0] == 'M' && blob[3] == 'V' && blob[6] == 2);
        CUDAVectorEngine engine2(VectorMetric::cosine, 4);
        engine2.deserialize(blob);
        EXPECT(contains(engine2.search({0, 1, 0, 0}, 10), 1) && engine2.count() == 2);
    }
    {   // metalVectorEngineAddBatchUpdatesExistingIdsCorrectly (:36-47)
        CUDAVectorEngine engine(VectorMetric::cosine, 2);
        engine.add(10, {1, 0});
        engine.add(20, {0, 1});
        engine.addBatch({20}, {{0.7f, 0.7f}});
        auto hits = engine.search({0.7f, 0.7f}, 1);
        EXPECT(hits.size() == 1 && hits[0].first == 20);
    }
    {   // error mapping: dimension mismatch -> EncodingError (MetalVectorEngine.swift:830-833)
        CUDAVectorEngine engine(VectorMetric::dot, 3);
        engine.add(1, {1, 2, 3});
        bool threw = false;
        try { engine.search({1, 2}, 1); } catch (const wax::EncodingError &) { threw = true; }
        EXPECT(threw);
        auto hits = engine.search({1, 2, 3}, 5);
        EXPECT(hits.size() == 1 && std::fabs(hits[0].second - (14.0f - 1.0f)) < 1e-5f);   // dot score = q.v - 1
    }
    std::printf("cpp mirror ok\n");
    return 0;
}
