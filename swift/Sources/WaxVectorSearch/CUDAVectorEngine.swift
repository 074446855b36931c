//
//  CUDAVectorEngine.swift
//
//  UNCOMPILED SOURCE: this image has no Swift toolchain (see DESIGN.md section 1). It is the literal
//  binding a Wax maintainer adds next to MetalVectorEngine.swift; it was desk-checked against
//  include/wax_vs_cuda.h but has never been through swiftc.
//
//  A `VectorSearchEngine` (Sources/WaxVectorSearch/VectorSearchEngine.swift:10-18) backed by
//  libwaxvs_cuda.so. Public surface mirrors `MetalVectorEngine` (MetalVectorEngine.swift:17):
//  isAvailable, init(metric:dimensions:), load(from:metric:dimensions:), search, add, addBatch,
//  addBatchStreaming, remove, serialize, deserialize, stageForCommit.
//
#if canImport(WaxVectorSearchCUDAC)
import Foundation
import WaxCore
import WaxVectorSearchCUDAC   // module map over include/wax_vs_cuda.h, linked against libwaxvs_cuda.so

public actor CUDAVectorEngine {
    private static let maxResults = 10_000

    private let metric: VectorMetric
    public let dimensions: Int
    private nonisolated(unsafe) let handle: OpaquePointer
    private let io: BlockingIOExecutor        // calls block the thread: never run them on the cooperative pool
    private var dirty = false

    public static var isAvailable: Bool {     // MetalVectorEngine.isAvailable (:144-146)
        var n: Int32 = 0
        return wax_vs_device_count(&n) == WAX_VS_OK && n > 0
    }

    public init(metric: VectorMetric, dimensions: Int) throws {
        guard dimensions > 0 else { throw WaxError.invalidToc(reason: "dimensions must be > 0") }
        guard dimensions <= Constants.maxEmbeddingDimensions else {
            throw WaxError.capacityExceeded(limit: UInt64(Constants.maxEmbeddingDimensions), requested: UInt64(dimensions))
        }
        var h: OpaquePointer?
        let rc = wax_vs_create(UInt32(dimensions), metric.toVecSimilarity().rawValue, nil, 0, &h)
        guard rc == WAX_VS_OK, let h else { throw Self.error(rc) }   // catchable: callers fall back to USearch (WaxSession.swift:484-497)
        self.handle = h
        self.metric = metric
        self.dimensions = dimensions
        self.io = BlockingIOExecutor(label: "com.wax.cuda", qos: .userInitiated)
    }

    deinit { wax_vs_destroy(handle) }

    /// MetalVectorEngine.load(from:metric:dimensions:) (:318-328)
    public static func load(from wax: Wax, metric: VectorMetric, dimensions: Int) async throws -> CUDAVectorEngine {
        let engine = try CUDAVectorEngine(metric: metric, dimensions: dimensions)
        if let bytes = try await wax.readCommittedVecIndexBytes() { try await engine.deserialize(bytes) }
        let pending = await wax.pendingEmbeddingMutations()
        if !pending.isEmpty {
            try await engine.addBatch(frameIds: pending.map(\.frameId), vectors: pending.map(\.vector))
        }
        return engine
    }

    public func search(vector: [Float], topK: Int) async throws -> [(frameId: UInt64, score: Float)] {
        let handle = self.handle
        let cap = min(max(topK, 1), Self.maxResults)
        return try await io.run {
            var ids = [UInt64](repeating: 0, count: cap)
            var scores = [Float](repeating: 0, count: cap)
            var n: UInt32 = 0
            let rc = vector.withUnsafeBufferPointer { q in
                wax_vs_search(handle, q.baseAddress, UInt32(vector.count), Int64(topK), &ids, &scores, UInt32(cap), &n)
            }
            guard rc == WAX_VS_OK else { throw Self.error(rc) }
            return (0..<Int(n)).map { (ids[$0], scores[$0]) }
        }
    }

    /// Batched form (no counterpart in the reference protocol): one tensor-core pass over the corpus nominates,
    /// an exact fp32 re-score decides; results are identical to calling `search` per query.
    public func searchBatch(vectors: [[Float]], topK: Int) async throws -> [[(frameId: UInt64, score: Float)]] {
        guard !vectors.isEmpty else { return [] }
        let dims = dimensions
        for v in vectors where v.count != dims {
            throw WaxError.encodingError(reason: "vector dimension mismatch: expected \(dims), got \(v.count)")
        }
        let handle = self.handle
        let cap = min(max(topK, 1), Self.maxResults)
        return try await io.run {
            var flat = [Float](); flat.reserveCapacity(vectors.count * dims)
            for v in vectors { flat.append(contentsOf: v) }
            var ids = [UInt64](repeating: 0, count: vectors.count * cap)
            var scores = [Float](repeating: 0, count: vectors.count * cap)
            var counts = [UInt32](repeating: 0, count: vectors.count)
            let rc = wax_vs_search_batch(handle, flat, UInt32(vectors.count), UInt32(dims), Int64(topK),
                                         &ids, &scores, UInt32(cap), &counts)
            guard rc == WAX_VS_OK else { throw Self.error(rc) }
            return (0..<vectors.count).map { q in (0..<Int(counts[q])).map { (ids[q * cap + $0], scores[q * cap + $0]) } }
        }
    }

    /// Frame filter pushed below the top-k (replaces the post-hoc filter + 3 x topK over-fetch of
    /// UnifiedSearch.swift:58,1241-1258): `allow == true` keeps only `frameIds`, `false` excludes them.
    public func search(vector: [Float], topK: Int, frameIds: [UInt64], allow: Bool) async throws -> [(frameId: UInt64, score: Float)] {
        let handle = self.handle
        let cap = min(max(topK, 1), Self.maxResults)
        return try await io.run {
            var ids = [UInt64](repeating: 0, count: cap)
            var scores = [Float](repeating: 0, count: cap)
            var n: UInt32 = 0
            let rc = wax_vs_search_filtered(handle, vector, UInt32(vector.count), Int64(topK), frameIds,
                                            UInt64(frameIds.count), allow ? 0 : 1, &ids, &scores, UInt32(cap), &n)
            guard rc == WAX_VS_OK else { throw Self.error(rc) }
            return (0..<Int(n)).map { (ids[$0], scores[$0]) }
        }
    }

    /// The same filter for a batch of queries in one pass over the corpus (wax_vs_search_batch_filtered).
    public func searchBatch(vectors: [[Float]], topK: Int, frameIds: [UInt64], allow: Bool) async throws -> [[(frameId: UInt64, score: Float)]] {
        guard !vectors.isEmpty else { return [] }
        let dims = dimensions
        for v in vectors where v.count != dims {
            throw WaxError.encodingError(reason: "vector dimension mismatch: expected \(dims), got \(v.count)")
        }
        let handle = self.handle
        let cap = min(max(topK, 1), Self.maxResults)
        return try await io.run {
            var flat = [Float](); flat.reserveCapacity(vectors.count * dims)
            for v in vectors { flat.append(contentsOf: v) }
            var ids = [UInt64](repeating: 0, count: vectors.count * cap)
            var scores = [Float](repeating: 0, count: vectors.count * cap)
            var counts = [UInt32](repeating: 0, count: vectors.count)
            let rc = wax_vs_search_batch_filtered(handle, flat, UInt32(vectors.count), UInt32(dims), Int64(topK), frameIds,
                                                  UInt64(frameIds.count), allow ? 0 : 1, &ids, &scores, UInt32(cap), &counts)
            guard rc == WAX_VS_OK else { throw Self.error(rc) }
            return (0..<vectors.count).map { q in (0..<Int(counts[q])).map { (ids[q * cap + $0], scores[q * cap + $0]) } }
        }
    }

    public func add(frameId: UInt64, vector: [Float]) async throws {
        try await addBatch(frameIds: [frameId], vectors: [vector])
    }

    public func addBatch(frameIds: [UInt64], vectors: [[Float]]) async throws {
        guard !frameIds.isEmpty else { return }
        guard frameIds.count == vectors.count else {
            throw WaxError.encodingError(reason: "addBatch: frameIds.count != vectors.count")
        }
        let dims = dimensions
        for v in vectors where v.count != dims {
            throw WaxError.encodingError(reason: "vector dimension mismatch: expected \(dims), got \(v.count)")
        }
        let handle = self.handle
        try await io.run {
            var flat = [Float](); flat.reserveCapacity(vectors.count * dims)
            for v in vectors { flat.append(contentsOf: v) }       // [[Float]] -> row-major n x dims
            let rc = wax_vs_add_batch(handle, frameIds, flat, UInt64(frameIds.count), UInt32(dims))
            guard rc == WAX_VS_OK else { throw Self.error(rc) }
        }
        dirty = true
    }

    public func addBatchStreaming(frameIds: [UInt64], vectors: [[Float]], chunkSize: Int = 256) async throws {
        guard frameIds.count > chunkSize else { return try await addBatch(frameIds: frameIds, vectors: vectors) }
        for start in stride(from: 0, to: frameIds.count, by: chunkSize) {
            let end = min(start + chunkSize, frameIds.count)
            try await addBatch(frameIds: Array(frameIds[start..<end]), vectors: Array(vectors[start..<end]))
        }
    }

    public func remove(frameId: UInt64) async throws {
        let handle = self.handle
        try await io.run {
            let rc = wax_vs_remove(handle, frameId)
            guard rc == WAX_VS_OK else { throw Self.error(rc) }
        }
        dirty = true
    }

    /// Many frames in one pass: one order-preserving compaction in HBM and one id-map rebuild instead of one
    /// tail `memmove` per id (`MetalVectorEngine.swift:431-441`).  Unknown ids are ignored.
    @discardableResult
    public func removeBatch(frameIds: [UInt64]) async throws -> Int {
        guard !frameIds.isEmpty else { return 0 }
        let handle = self.handle
        let removed: UInt64 = try await io.run {
            var gone: UInt64 = 0
            let rc = frameIds.withUnsafeBufferPointer { wax_vs_remove_batch(handle, $0.baseAddress, UInt64($0.count), &gone) }
            guard rc == WAX_VS_OK else { throw Self.error(rc) }
            return gone
        }
        if removed > 0 { dirty = true }
        return Int(removed)
    }

    public func serialize() async throws -> Data {
        let handle = self.handle
        return try await io.run {
            var len: UInt64 = 0
            guard wax_vs_serialized_length(handle, &len) == WAX_VS_OK else { throw Self.error(WAX_VS_ERR_CUDA) }
            var data = Data(count: Int(len))
            let rc = data.withUnsafeMutableBytes { wax_vs_serialize(handle, $0.bindMemory(to: UInt8.self).baseAddress, len, &len) }
            guard rc == WAX_VS_OK else { throw Self.error(rc) }
            return data
        }
    }

    public func deserialize(_ data: Data) async throws {
        let handle = self.handle
        try await io.run {
            let rc = data.withUnsafeBytes { wax_vs_deserialize(handle, $0.bindMemory(to: UInt8.self).baseAddress, UInt64(data.count)) }
            guard rc == WAX_VS_OK else { throw Self.error(rc) }
        }
        dirty = false
    }

    public func stageForCommit(into wax: Wax) async throws {      // MetalVectorEngine.swift:818-828
        if !dirty { return }
        let blob = try await serialize()
        var count: UInt64 = 0
        _ = wax_vs_count(handle, &count)
        try await wax.stageVecIndexForNextCommit(bytes: blob, vectorCount: count, dimension: UInt32(dimensions),
                                                 similarity: metric.toVecSimilarity())
        dirty = false
    }

    /// rc -> the WaxError case the Metal engine throws for the same condition.
    private static func error(_ rc: Int32) -> WaxError {
        let reason = String(cString: wax_vs_last_error())
        switch rc {
        case WAX_VS_ERR_DIMENSION: return .encodingError(reason: reason)
        case WAX_VS_ERR_CAPACITY: return .capacityExceeded(limit: UInt64(UInt32.max), requested: 0)
        default: return .invalidToc(reason: reason)
        }
    }
}

extension CUDAVectorEngine: VectorSearchEngine {}
#endif
