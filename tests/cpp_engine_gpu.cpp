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
    {   // searchBatch == one search per query (130 queries x 64 dims: the tensor-core levels), reserve, streaming add
        CUDAVectorEngine engine(VectorMetric::cosine, 64);
        engine.reserve(3000);
        std::vector<uint64_t> ids;
        std::vector<std::vector<float>> rows;
        uint32_t state = 12345u;
        auto rnd = [&]() { state = state * 1664525u + 1013904223u; return static_cast<float>(state >> 8) / 8388608.0f - 1.0f; };
        for (uint64_t i = 0; i < 3000; ++i) {
            std::vector<float> v(64);
            for (auto &x : v) x = rnd();
            ids.push_back(1000 + i);
            rows.push_back(v);
        }
        engine.addBatchStreaming(ids, rows, 700);
        EXPECT(engine.count() == 3000);
        std::vector<std::vector<float>> qs(rows.begin() + 5, rows.begin() + 135);
        auto batch = engine.searchBatch(qs, 10);
        EXPECT(batch.size() == 130);
        for (size_t q : {size_t(0), size_t(64), size_t(129)}) {
            auto one = engine.search(qs[q], 10);
            EXPECT(batch[q].size() == 10 && one.size() == 10 && batch[q][0].first == 1005 + q);
            for (size_t i = 0; i < 10; ++i) EXPECT(batch[q][i].first == one[i].first && batch[q][i].second == one[i].second);
        }
        // filter pushed below the top-k (UnifiedSearchTests.swift:133-158: allow-list {id2, id3} with topK 2)
        auto allowed = engine.searchFiltered(qs[0], 2, {1002, 1003}, true);
        EXPECT(allowed.size() == 2 && (allowed[0].first == 1002 || allowed[0].first == 1003));
        auto denied = engine.searchFiltered(qs[0], 3, {1005}, false);
        EXPECT(denied.size() == 3 && !contains(denied, 1005));
        // ... and its batched form: one filter, every query, the same answers as the per-query calls
        std::vector<uint64_t> deny;
        for (uint64_t i = 0; i < 2000; ++i) deny.push_back(1000 + i);
        auto fb = engine.searchBatchFiltered(qs, 5, deny, false);
        EXPECT(fb.size() == 130);
        for (size_t q : {size_t(0), size_t(77), size_t(129)}) {
            auto one = engine.searchFiltered(qs[q], 5, deny, false);
            EXPECT(fb[q].size() == 5 && one.size() == 5);
            for (size_t i = 0; i < 5; ++i) EXPECT(fb[q][i].first == one[i].first && fb[q][i].second == one[i].second && fb[q][i].first >= 3000);
        }
        auto fa = engine.searchBatchFiltered(qs, 2, {1002, 1003}, true);
        EXPECT(fa.size() == 130 && fa[5].size() == 2 && (fa[5][0].first == 1002 || fa[5][0].first == 1003));
    }
    {   // load(from:): committed blob + pending embeddings replayed as upserts (MetalVectorEngine.swift:318-328)
        CUDAVectorEngine src(VectorMetric::cosine, 4);
        src.addBatch({0, 1, 2}, {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}});
        auto blob = src.serialize();
        CUDAVectorEngine *loaded = CUDAVectorEngine::load(&blob, {3, 1, 3}, {{0, 0, 0, 1}, {0.6f, 0.8f, 0, 0}, {0, 0, 0.6f, 0.8f}},
                                                          VectorMetric::cosine, 4);
        EXPECT(loaded->count() == 4);
        EXPECT(loaded->search({0, 0, 0.6f, 0.8f}, 1)[0].first == 3 && loaded->search({0.6f, 0.8f, 0, 0}, 1)[0].first == 1);
        delete loaded;
        CUDAVectorEngine *empty = CUDAVectorEngine::load(nullptr, {}, {}, VectorMetric::cosine, 4);
        EXPECT(empty->count() == 0 && empty->search({1, 0, 0, 0}, 3).empty());
        delete empty;
    }
    std::printf("cpp mirror ok\n");
    return 0;
}
