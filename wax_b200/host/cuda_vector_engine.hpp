// cuda_vector_engine.hpp -- header-only C++17 mirror of Wax's `VectorSearchEngine` over the C-ABI
// (for C++ hosts; the Swift actor in swift/ and the Python mirror in wax_b200/engine.py bind the same entry
// points).  Member names follow the protocol (Sources/WaxVectorSearch/VectorSearchEngine.swift:10-18) and
// MetalVectorEngine's public surface (MetalVectorEngine.swift:144-146,153,330-446,682-815).
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/wax_vs_cuda.h"

namespace wax {

enum class VectorMetric : uint8_t { cosine = WAX_VS_COSINE, dot = WAX_VS_DOT, l2 = WAX_VS_L2 };

// WaxError cases thrown on this path.
struct WaxError : std::runtime_error { using std::runtime_error::runtime_error; };
struct EncodingError : WaxError { using WaxError::WaxError; };
struct CapacityExceeded : WaxError { using WaxError::WaxError; };
struct InvalidToc : WaxError { using WaxError::WaxError; };

class CUDAVectorEngine {
public:
    using Hit = std::pair<uint64_t, float>;  // (frameId, score)

    static bool isAvailable() {
        int32_t n = 0;
        return wax_vs_device_count(&n) == WAX_VS_OK && n > 0;
    }
    CUDAVectorEngine(VectorMetric metric, uint32_t dimensions) : metric_(metric), dimensions_(dimensions) {
        check(wax_vs_create(dimensions, static_cast<uint8_t>(metric), nullptr, 0, &h_));
    }
    ~CUDAVectorEngine() { wax_vs_destroy(h_); }
    CUDAVectorEngine(const CUDAVectorEngine &) = delete;
    CUDAVectorEngine &operator=(const CUDAVectorEngine &) = delete;

    uint32_t dimensions() const { return dimensions_; }
    uint64_t count() const { uint64_t n = 0; check(wax_vs_count(h_, &n)); return n; }

    std::vector<Hit> search(const std::vector<float> &vector, int64_t topK) const {
        const int64_t lim = topK < 1 ? 1 : (topK > WAX_VS_MAX_RESULTS ? WAX_VS_MAX_RESULTS : topK);
        std::vector<uint64_t> ids(static_cast<size_t>(lim));
        std::vector<float> scores(static_cast<size_t>(lim));
        uint32_t n = 0;
        check(wax_vs_search(h_, vector.data(), static_cast<uint32_t>(vector.size()), topK, ids.data(), scores.data(),
                            static_cast<uint32_t>(lim), &n));
        std::vector<Hit> out(n);
        for (uint32_t i = 0; i < n; ++i) out[i] = {ids[i], scores[i]};
        return out;
    }
    void add(uint64_t frameId, const std::vector<float> &vector) {
        check(wax_vs_add(h_, frameId, vector.data(), static_cast<uint32_t>(vector.size())));
    }
    void addBatch(const std::vector<uint64_t> &frameIds, const std::vector<std::vector<float>> &vectors) {
        if (frameIds.empty()) return;
        if (frameIds.size() != vectors.size()) throw EncodingError("addBatch: frameIds.count != vectors.count");
        std::vector<float> flat;
        flat.reserve(vectors.size() * dimensions_);
        for (const auto &v : vectors) {
            if (v.size() != dimensions_)
                throw EncodingError("vector dimension mismatch: expected " + std::to_string(dimensions_) + ", got " +
                                    std::to_string(v.size()));
            flat.insert(flat.end(), v.begin(), v.end());
        }
        check(wax_vs_add_batch(h_, frameIds.data(), flat.data(), frameIds.size(), dimensions_));
    }
    // addBatchStreaming (MetalVectorEngine.swift:404-421): chunks of `chunkSize` through addBatch.
    void addBatchStreaming(const std::vector<uint64_t> &frameIds, const std::vector<std::vector<float>> &vectors,
                           size_t chunkSize = 256) {
        if (frameIds.empty()) return;
        if (frameIds.size() != vectors.size()) throw EncodingError("addBatchStreaming: frameIds.count != vectors.count");
        if (chunkSize == 0) chunkSize = 256;
        for (size_t lo = 0; lo < frameIds.size(); lo += chunkSize) {
            const size_t hi = lo + chunkSize < frameIds.size() ? lo + chunkSize : frameIds.size();
            addBatch(std::vector<uint64_t>(frameIds.begin() + lo, frameIds.begin() + hi),
                     std::vector<std::vector<float>>(vectors.begin() + lo, vectors.begin() + hi));
        }
    }
    void reserve(uint64_t rows) { check(wax_vs_reserve(h_, rows)); }   // reserveIfNeeded (:857-871)
    void remove(uint64_t frameId) { check(wax_vs_remove(h_, frameId)); }
    // Many frames in one pass (one compaction in HBM, one hash rebuild); returns how many rows were deleted.
    uint64_t removeBatch(const std::vector<uint64_t> &frameIds) {
        uint64_t gone = 0;
        check(wax_vs_remove_batch(h_, frameIds.data(), frameIds.size(), &gone));
        return gone;
    }

    // A batch of independent queries (no reference counterpart: VectorSearchEngine.swift:13 takes one vector).  Eligible
    // batches take the tensor-core levels; the results are identical to one search() per query.
    std::vector<std::vector<Hit>> searchBatch(const std::vector<std::vector<float>> &vectors, int64_t topK) const {
        std::vector<std::vector<Hit>> out(vectors.size());
        if (vectors.empty()) return out;
        std::vector<float> flat;
        flat.reserve(vectors.size() * dimensions_);
        for (const auto &v : vectors) {
            if (v.size() != dimensions_)
                throw EncodingError("vector dimension mismatch: expected " + std::to_string(dimensions_) + ", got " +
                                    std::to_string(v.size()));
            flat.insert(flat.end(), v.begin(), v.end());
        }
        const int64_t lim = topK < 1 ? 1 : (topK > WAX_VS_MAX_RESULTS ? WAX_VS_MAX_RESULTS : topK);
        // stride = min(clamp(topK), count); a concurrent add may grow the count between count() and the search, the
        // library then answers WAX_VS_ERR_BUFFER and the buffers are re-sized (clamp(topK) always suffices)
        const uint64_t rows = count();
        uint32_t stride = static_cast<uint32_t>(rows < static_cast<uint64_t>(lim) ? (rows ? rows : 1) : lim);
        std::vector<uint64_t> ids;
        std::vector<float> scores;
        std::vector<uint32_t> ns(vectors.size());
        int32_t rc = WAX_VS_OK;
        for (int attempt = 0; attempt < 2; ++attempt) {
            ids.assign(vectors.size() * stride, 0);
            scores.assign(vectors.size() * stride, 0.0f);
            rc = wax_vs_search_batch(h_, flat.data(), static_cast<uint32_t>(vectors.size()), dimensions_, topK, ids.data(),
                                     scores.data(), stride, ns.data());
            if (rc != WAX_VS_ERR_BUFFER) break;
            stride = static_cast<uint32_t>(lim);
        }
        check(rc);
        for (size_t q = 0; q < vectors.size(); ++q) {
            out[q].resize(ns[q]);
            for (uint32_t i = 0; i < ns[q]; ++i) out[q][i] = {ids[q * stride + i], scores[q * stride + i]};
        }
        return out;
    }

    // The frame filter of UnifiedSearch (UnifiedSearch.swift:58,1195-1200,1241-1258) pushed below the top-k:
    // allow == true: only the listed frameIds may be returned; false: they are excluded (deleted / superseded frames).
    std::vector<Hit> searchFiltered(const std::vector<float> &vector, int64_t topK, const std::vector<uint64_t> &frameIds,
                                    bool allow) const {
        const int64_t lim = topK < 1 ? 1 : (topK > WAX_VS_MAX_RESULTS ? WAX_VS_MAX_RESULTS : topK);
        std::vector<uint64_t> ids(static_cast<size_t>(lim));
        std::vector<float> scores(static_cast<size_t>(lim));
        uint32_t n = 0;
        check(wax_vs_search_filtered(h_, vector.data(), static_cast<uint32_t>(vector.size()), topK, frameIds.data(),
                                     frameIds.size(), allow ? 0 : 1, ids.data(), scores.data(), static_cast<uint32_t>(lim), &n));
        std::vector<Hit> out(n);
        for (uint32_t i = 0; i < n; ++i) out[i] = {ids[i], scores[i]};
        return out;
    }

    // searchFiltered for a batch of queries under ONE filter, one pass over the corpus (wax_vs_search_batch_filtered).
    std::vector<std::vector<Hit>> searchBatchFiltered(const std::vector<std::vector<float>> &vectors, int64_t topK,
                                                      const std::vector<uint64_t> &frameIds, bool allow) const {
        std::vector<std::vector<Hit>> out(vectors.size());
        if (vectors.empty()) return out;
        std::vector<float> flat;
        flat.reserve(vectors.size() * dimensions_);
        for (const auto &v : vectors) {
            if (v.size() != dimensions_)
                throw EncodingError("vector dimension mismatch: expected " + std::to_string(dimensions_) + ", got " +
                                    std::to_string(v.size()));
            flat.insert(flat.end(), v.begin(), v.end());
        }
        const uint32_t lim = static_cast<uint32_t>(topK < 1 ? 1 : (topK > WAX_VS_MAX_RESULTS ? WAX_VS_MAX_RESULTS : topK));
        std::vector<uint64_t> ids(vectors.size() * lim);
        std::vector<float> scores(vectors.size() * lim);
        std::vector<uint32_t> ns(vectors.size());
        check(wax_vs_search_batch_filtered(h_, flat.data(), static_cast<uint32_t>(vectors.size()), dimensions_, topK,
                                           frameIds.data(), frameIds.size(), allow ? 0 : 1, ids.data(), scores.data(), lim,
                                           ns.data()));
        for (size_t q = 0; q < vectors.size(); ++q) {
            out[q].resize(ns[q]);
            for (uint32_t i = 0; i < ns[q]; ++i) out[q][i] = {ids[q * lim + i], scores[q * lim + i]};
        }
        return out;
    }

    // static load(from:metric:dimensions:) (MetalVectorEngine.swift:318-328): the committed blob (may be empty = none
    // committed yet), then the pending embedding mutations as ONE upsert batch (sequential semantics in the library).
    static CUDAVectorEngine *load(const std::vector<uint8_t> *committedBlob, const std::vector<uint64_t> &pendingIds,
                                  const std::vector<std::vector<float>> &pendingVectors, VectorMetric metric,
                                  uint32_t dimensions) {
        auto *engine = new CUDAVectorEngine(metric, dimensions);
        try {
            if (committedBlob) engine->deserialize(*committedBlob);
            engine->addBatch(pendingIds, pendingVectors);
        } catch (...) {
            delete engine;
            throw;
        }
        return engine;
    }
    std::vector<uint8_t> serialize() const {
        uint64_t len = 0;
        check(wax_vs_serialized_length(h_, &len));
        std::vector<uint8_t> blob(len);
        check(wax_vs_serialize(h_, blob.data(), len, &len));
        return blob;
    }
    void deserialize(const std::vector<uint8_t> &blob) { check(wax_vs_deserialize(h_, blob.data(), blob.size())); }
    wax_vs_engine *handle() const { return h_; }

private:
    static void check(int32_t rc) {
        if (rc == WAX_VS_OK) return;
        const std::string reason = wax_vs_last_error();
        if (rc == WAX_VS_ERR_DIMENSION) throw EncodingError(reason);
        if (rc == WAX_VS_ERR_CAPACITY) throw CapacityExceeded(reason);
        throw InvalidToc(reason);
    }
    VectorMetric metric_;
    uint32_t dimensions_;
    wax_vs_engine *h_ = nullptr;
};

}  // namespace wax
