// mpp::TPS_Astar (product code): A* over the SE(2) lattice with the per-node collision rows and the
// per-edge evaluator costs computed by the GPU library in batches.
//
// Reference: src/algos/TPS_Astar.cpp (plan :86-475, find_feasible_paths_to_neighbors :538-826,
// heuristics :477-537) and include/mpp/algos/TPS_Astar.h. The reference interleaves, per expanded
// node, candidate enumeration, collision checks, neighbour selection, edge construction, cost
// evaluation and relaxation. Here the same work is cut into phases so that the two data-parallel
// parts leave the host:
//   A  (host)  enumerate the candidate TPS points of the node, reconstruct their end poses
//   B  (GPU)   free TP-distance of every (node, PTG, k[, speed])      -> mppb_eval_nodes / _holo_
//   C  (host)  blocked test, goal-cell rule, best candidate per reached lattice cell
//   D  (host)  neighbour nodes, tentative edges and their interpolated samples
//   E  (GPU)   evaluator costs of all tentative edges                  -> mppb_eval_edge_costs
//   F  (host)  relaxation: gScore/fScore, open set, motion tree
// Every container whose iteration order decides results (the unordered_map of best paths, the
// multimap open set with its stale keys, node-id assignment order) is kept as in the reference,
// and every PTG object sees the same sequence of state changes, so a query produces the same tree
// whether it runs alone (plan) or interleaved with thousands of others (plan_batch).
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <tuple>
#include <map>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <memory_resource>
#include <cstring>
#include <mutex>
#include <set>
#include <sstream>
#include <thread>
#include <unordered_map>
#include <unordered_set>

#include "mpp.h"

namespace mpp
{
namespace
{
#ifdef MPPB_PROFILE  // host-phase attribution experiments only (build with MPPB_EXTRA_NVCC=-DMPPB_PROFILE)
#include <x86intrin.h>
struct ProfCounters
{
    std::atomic<uint64_t> t[8] = {}, n[8] = {};
    ~ProfCounters()
    {
        static const char* names[8] = {"pop", "A", "order+prefetch", "find+prefetch", "relax loop", "heuristic", "heap insert", "-"};
        for (int i = 0; i < 8; i++)
            if (n[i]) fprintf(stderr, "[mppb prof] %-16s %12llu calls %8.1f cycles/call %10.3f Gcycles\n", names[i],
                              (unsigned long long)n[i].load(), double(t[i].load()) / double(n[i].load()), double(t[i].load()) * 1e-9);
    }
};
static ProfCounters g_prof;
struct ProfScope
{
    int      i;
    uint64_t t0;
    explicit ProfScope(int i_) : i(i_), t0(__rdtsc()) {}
    ~ProfScope()
    {
        g_prof.t[i].fetch_add(__rdtsc() - t0, std::memory_order_relaxed);
        g_prof.n[i].fetch_add(1, std::memory_order_relaxed);
    }
};
#define MPPB_PROF(i) ProfScope profScope##i(i)
#else
#define MPPB_PROF(i)
#endif
// a known neighbour's second cache line (the state an accepted relaxation overwrites) is prefetched with its first:
// C4, 4096 queries at cap 2000 on 16 host threads: 5.9-6.2 s -> 4.9-5.6 s (MPPB_NO_PF2=1 for A/B measurements)
const bool kPrefetchStateLine = std::getenv("MPPB_NO_PF2") == nullptr;
// batches: a worker requests the hot fields, neighbour records and current node of the queries it will turn to next
// (C4, 4096 queries at cap 2000 on 16 host threads: 5.1 s -> 4.5-4.8 s; MPPB_NO_PFQ=1 for A/B measurements)
const bool kPrefetchNextQuery = std::getenv("MPPB_NO_PFQ") == nullptr;
using Clock = std::chrono::steady_clock;
double seconds_since(const Clock::time_point& t0) { return std::chrono::duration<double>(Clock::now() - t0).count(); }

void ck(mppb_ctx* ctx, int rc)
{
    if (rc != MPPB_OK) throw std::runtime_error(std::string("libmpp_b200: ") + mppb_last_error(ctx));
}

// ---- lattice ---------------------------------------------------------------------------------------
struct NodeCoords
{
    int32_t                idxX = 0, idxY = 0;
    std::optional<int32_t> idxYaw;
    NodeCoords() = default;
    NodeCoords(int32_t ix, int32_t iy) : idxX(ix), idxY(iy) {}
    NodeCoords(int32_t ix, int32_t iy, int32_t iphi) : idxX(ix), idxY(iy), idxYaw(iphi) {}
    bool operator==(const NodeCoords& o) const { return idxX == o.idxX && idxY == o.idxY && idxYaw == o.idxYaw; }
    /** same (x,y) cell and, when both carry a heading index, the same heading */
    bool sameLocation(const NodeCoords& o) const
    {
        if (idxX != o.idxX || idxY != o.idxY) return false;
        return !(idxYaw.has_value() && o.idxYaw.has_value()) || *idxYaw == *o.idxYaw;
    }
};
struct NodeCoordsHash
{
    size_t operator()(const NodeCoords& x) const
    {
        size_t res = 17;
        res        = res * 31 + std::hash<int32_t>()(x.idxX);
        res        = res * 31 + std::hash<int32_t>()(x.idxY);
        if (x.idxYaw) res = res * 31 + std::hash<int32_t>()(x.idxYaw.value());
        return res;
    }
};
// One lattice node. With thousands of interleaved queries the search state is far larger than the caches and every
// neighbour of an expansion is a cache miss, so the fields are ordered by use: the first 64-byte line holds everything
// the common case reads (a neighbour that exists already and is not improved: id, visited, gScore), the second what an
// accepted relaxation overwrites, the third what only the tree builder reads.
struct alignas(64) Node
{
    // ---- line 1 ----
    std::optional<TNodeID> id;
    cost_t                 gScore = std::numeric_limits<cost_t>::max();
    cost_t                 fScore = std::numeric_limits<cost_t>::max();
    const Node*            cameFrom         = nullptr;
    bool                   pendingInOpenSet = false;
    bool                   visited          = false;
    // Device-expansion path: the node's place in the motion tree, kept here instead of in a tree that is mutated
    // while searching. A node is expanded at most once (visited nodes are never relaxed again), so when a relaxation
    // is accepted its parent's cost and state are final: the node's tree cost IS its gScore, its edge is described by
    // the edge* fields, and a parent's edge list is its surviving children in the order of their accepting
    // relaxations (edgeSeq). The reference's rewire_node_parent leaves the node's own state as first inserted
    // (MotionPrimitivesTree.h:173-215): firstState.
    bool     everParent = false;  // some relaxation out of this node was accepted (it owns an edge list)
    uint8_t  edgePtg    = 0;
    uint16_t edgeK      = 0;
    uint32_t edgeStep   = 0;
    uint32_t edgeSeq    = 0;  // number of the accepting relaxation within the query
    float    edgeDist   = 0;
    // ---- line 2 ----
    alignas(64) SE2_KinState state;
    cost_t edgeCost = 0;
    // ---- line 3 ----
    alignas(64) SE2_KinState firstState;
};
static_assert(sizeof(Node) == 192 && offsetof(Node, state) == 64 && offsetof(Node, firstState) == 128, "Node layout");
// The SE(2) lattice of the reference is a std::unordered_map<NodeCoords, Node> that is only ever indexed, never
// iterated (TPS_Astar.h:222, TPS_Astar.cpp:514-525), so its container type cannot influence the result. Here: an
// open-addressing index (linear probing, load <= 1/2) over nodes that live in the query's arena (stable addresses:
// the open set and cameFrom hold Node pointers).
class LatticeIndex
{
   public:
    explicit LatticeIndex(std::pmr::monotonic_buffer_resource* arena) : arena_(arena) { slots_.assign(1024, Slot{}); }
    Node& operator[](const NodeCoords& c)
    {
        const int32_t yaw = c.idxYaw.value();  // lattice nodes always carry a heading index
        for (;;)
        {
            const size_t mask = slots_.size() - 1;
            size_t       h    = mix(c.idxX, c.idxY, yaw) & mask;
            for (;; h = (h + 1) & mask)
            {
                Slot& s = slots_[h];
                if (!s.node) break;
                if (s.x == c.idxX && s.y == c.idxY && s.yaw == yaw) return *s.node;
            }
            if ((count_ + 1) * 2 > slots_.size())
            {
                grow();
                continue;
            }
            Slot& s = slots_[h];
            s.x = c.idxX, s.y = c.idxY, s.yaw = yaw;
            s.node = new (arena_->allocate(sizeof(Node), alignof(Node))) Node();
            count_++;
            return *s.node;
        }
    }

    /** the node of (x, y, yaw) if it exists, else nullptr (nothing is created) */
    Node* find(int32_t x, int32_t y, int32_t yaw) const
    {
        const size_t mask = slots_.size() - 1;
        for (size_t h = mix(x, y, yaw) & mask;; h = (h + 1) & mask)
        {
            const Slot& s = slots_[h];
            if (!s.node) return nullptr;
            if (s.x == x && s.y == y && s.yaw == yaw) return s.node;
        }
    }
    /** every node of the lattice (any order) */
    template <class F>
    void forEachNode(F&& f) const
    {
        for (const Slot& s : slots_)
            if (s.node) f(*s.node);
    }
    /** warm the cache line the look-up of (x, y, yaw) will start at (the search state of many interleaved queries is
     *  far larger than the caches) */
    void prefetch(int32_t x, int32_t y, int32_t yaw) const
    {
        __builtin_prefetch(&slots_[mix(x, y, yaw) & (slots_.size() - 1)]);
    }

   private:
    struct Slot
    {
        int32_t x = 0, y = 0, yaw = 0;
        Node*   node = nullptr;
    };
    static size_t mix(int32_t x, int32_t y, int32_t yaw)
    {
        uint64_t h = static_cast<uint32_t>(x) * 0x9E3779B97F4A7C15ull;
        h ^= (static_cast<uint64_t>(static_cast<uint32_t>(y)) + 0x7F4A7C15ull) * 0xC2B2AE3D27D4EB4Full;
        h ^= (static_cast<uint64_t>(static_cast<uint32_t>(yaw)) + 0x165667B1ull) * 0xD6E8FEB86659FD93ull;
        h ^= h >> 29;
        return static_cast<size_t>(h * 0x9E3779B97F4A7C15ull >> 20);
    }
    void grow()
    {
        std::vector<Slot> old(slots_.size() * 2);
        old.swap(slots_);
        const size_t mask = slots_.size() - 1;
        for (const Slot& s : old)
        {
            if (!s.node) continue;
            size_t h = mix(s.x, s.y, s.yaw) & mask;
            while (slots_[h].node) h = (h + 1) & mask;
            slots_[h] = s;
        }
    }
    std::pmr::monotonic_buffer_resource* arena_;
    std::vector<Slot>                    slots_;
    size_t                               count_ = 0;
};
// The order in which the reference visits the cells an expansion reaches is the iteration order of its
// std::unordered_map<NodeCoords, path_to_neighbor_t> (TPS_Astar.cpp:565-566, 806-814), which assigns the node ids. On
// the device-expansion path the cells arrive distinct and in first-insertion order, so all the host needs from that
// container is the resulting ORDER. CellOrder replays libstdc++'s hashtable for exactly that — a singly linked list
// threaded through the buckets, a new node going to the front of its bucket or, if the bucket is empty, to the front of
// the whole list; a rehash re-threads the nodes in list order (_M_insert_bucket_begin / _M_rehash_aux, unique keys) —
// on index arrays, without allocation, look-up or node construction. The bucket counts come from a real
// std::unordered_map of this very library, grown once at start-up. selftest_cell_order() compares it with the real
// container on random key sets (tests/test_host_cpu.py).
class CellOrder
{
   public:
    static constexpr uint16_t kNil = 0xffff;
    void reset()
    {
        n_       = 0;
        head_    = kNil;
        level_   = 0;
        buckets_ = growth().counts[0];
        std::fill(bucketBefore_.begin(), bucketBefore_.begin() + static_cast<std::ptrdiff_t>(std::min(bucketBefore_.size(), buckets_)), kEmpty);
    }
    /** the n-th key (n = number of keys so far) with hash code `code`; keys must be distinct */
    void insert(size_t code)
    {
        const Growth& G = growth();
        if (level_ + 1 < G.counts.size() && n_ + 1 > G.maxSize[level_]) rehash(G.counts[++level_]);
        if (n_ >= next_.size())
        {
            next_.resize(std::max<size_t>(64, 2 * next_.size()));
            code_.resize(next_.size());
        }
        const uint16_t me = static_cast<uint16_t>(n_++);
        code_[me]         = code;
        link(me, code % buckets_);
    }
    uint16_t first() const { return head_; }
    uint16_t next(uint16_t i) const { return next_[i]; }

   private:
    // bucketBefore_[b]: the list node BEFORE the first node of bucket b; kHead = the list head itself; kEmpty = no node
    static constexpr uint16_t kEmpty = 0xfffe, kHead = 0xfffd;
    struct Growth
    {
        std::vector<size_t> counts;   // bucket counts the container steps through, from its first allocation
        std::vector<size_t> maxSize;  // elements it holds at most with counts[i] buckets before it rehashes
    };
    static const Growth& growth()
    {
        static const Growth g = [] {
            Growth                       r;
            std::unordered_map<int, int> m;
            for (int i = 0; i < 20000; i++)
            {
                m[i];
                if (r.counts.empty() || m.bucket_count() != r.counts.back())
                {
                    if (!r.counts.empty()) r.maxSize.push_back(static_cast<size_t>(i));
                    r.counts.push_back(m.bucket_count());
                }
            }
            r.maxSize.push_back(static_cast<size_t>(-1));
            return r;
        }();
        return g;
    }
    void link(uint16_t me, size_t b)
    {
        if (b >= bucketBefore_.size()) bucketBefore_.resize(b + 1, kEmpty);
        const uint16_t before = bucketBefore_[b];
        if (before != kEmpty)
        {
            // bucket not empty: the new node goes right after the node before the bucket's first one
            uint16_t& link = before == kHead ? head_ : next_[before];
            next_[me]      = link;
            link           = me;
        }
        else
        {
            // empty bucket: front of the whole list; the former first node's bucket now starts after the new node
            next_[me] = head_;
            head_     = me;
            if (next_[me] != kNil) bucketBefore_[code_[next_[me]] % buckets_] = me;
            bucketBefore_[b] = kHead;
        }
    }
    void rehash(size_t newCount)
    {
        buckets_ = newCount;
        if (bucketBefore_.size() < newCount) bucketBefore_.resize(newCount);
        std::fill(bucketBefore_.begin(), bucketBefore_.begin() + static_cast<std::ptrdiff_t>(newCount), kEmpty);
        uint16_t p = head_;
        head_      = kNil;
        size_t bbeginBkt = 0;
        while (p != kNil)
        {
            const uint16_t nxt = next_[p];
            const size_t   b   = code_[p] % newCount;
            if (bucketBefore_[b] == kEmpty)
            {
                next_[p]         = head_;
                head_            = p;
                bucketBefore_[b] = kHead;
                if (next_[p] != kNil) bucketBefore_[bbeginBkt] = p;
                bbeginBkt = b;
            }
            else
            {
                const uint16_t before = bucketBefore_[b];
                uint16_t&      link   = before == kHead ? head_ : next_[before];
                next_[p]              = link;
                link                  = p;
            }
            p = nxt;
        }
    }
    std::vector<uint16_t> next_, bucketBefore_;
    std::vector<size_t>   code_;
    size_t                n_ = 0, buckets_ = 1, level_ = 0;
    uint16_t              head_ = kNil;
};

struct PathToNeighbor
{
    std::optional<size_t>    ptgIndex;
    trajectory_index_t       ptgTrajIndex = -1;
    ptg_step_t               relTrgStep   = 0;
    TNavDynamicState         ptgDynState;
    normalized_speed_t       ptgTrimmableSpeed = 1.0;  // never assigned by the reference: stays 1.0
    distance_t               ptgDist           = std::numeric_limits<distance_t>::max();
    TPose2D                  relReconstrPose;
};
// one candidate TPS point whose end pose passed the world-box test (phase A output)
struct Candidate
{
    size_t             ptgIdx;
    trajectory_index_t k;
    ptg_step_t         step;
    distance_t         relTrgDist;
    TPose2D            relPose;
    NodeCoords         nc;
    distance_t         initDist;   // initTPObstacleSingle(k) at the moment the reference would take it
    bool               isTarget;   // direct-to-goal TPS point
    bool               onGrid;     // result lives in the collision-grid row (else in the holo list)
    uint32_t           evalIndex;  // column of the row / index into the query's holo candidates
    uint32_t           copies;     // how many identical TPS points of the reference this candidate stands for
};
struct PendingEdge
{
    Node*           neighbor;
    SE2_KinState    x_i;
    MoveEdgeSE2_TPS edge;
};

// The reference's open set is a std::multimap<cost, Node*> that is only ever inserted into and popped at begin()
// (keys go stale, nothing is erased elsewhere). A binary heap ordered by (key, insertion number) pops in exactly the
// multimap's order — equal keys leave in insertion order, as multimap::insert places them at the upper bound of
// their range — without a node allocation and a cold pointer chase per insertion.
// (selftest_open_set() below checks the equivalence against a std::multimap; tests/test_host_cpu.py runs it.)
template <class T>
struct OpenSetT
{
    struct Item
    {
        distance_t f;
        uint64_t   seq;
        T*         node;
    };
    struct Later
    {
        bool operator()(const Item& a, const Item& b) const { return a.f > b.f || (a.f == b.f && a.seq > b.seq); }
    };
    std::vector<Item> heap;
    uint64_t          nextSeq = 0;
    bool              empty() const { return heap.empty(); }
    T*                top() const { return heap.front().node; }
    void              insert(distance_t f, T* n)
    {
        heap.push_back(Item{f, nextSeq++, n});
        std::push_heap(heap.begin(), heap.end(), Later());
    }
    void pop()
    {
        std::pop_heap(heap.begin(), heap.end(), Later());
        heap.pop_back();
    }
};

// ---- persistent host thread pool for the per-query phases --------------------------------------------
// Items are handed out with affinity: worker t owns the t-th contiguous slice of [0,count) and only helps with the
// slices of others once its own is done. The planner runs the same per-query phases round after round, so a query's
// search state (lattice, open set, PTG clones, arenas) stays in the caches of one core instead of moving to whichever
// thread happened to ask next.
class Workers
{
   public:
    explicit Workers(int n) : n_(std::max(1, n)), slices_(static_cast<size_t>(std::max(1, n)))
    {
        for (int i = 1; i < n_; i++) threads_.emplace_back([this, i]() { loop(i); });
    }
    ~Workers()
    {
        {
            std::lock_guard<std::mutex> lk(m_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto& t : threads_) t.join();
    }
    int size() const { return n_; }
    // fn(i) for i in [0,count); exceptions are collected and the first one rethrown
    void run(size_t count, const std::function<void(size_t)>& fn)
    {
        if (count == 0) return;
        if (n_ == 1 || count == 1)
        {
            for (size_t i = 0; i < count; i++) fn(i);
            return;
        }
        {
            std::lock_guard<std::mutex> lk(m_);
            fn_      = &fn;
            pending_ = n_ - 1;
            err_.clear();
            for (int t = 0; t < n_; t++)
            {
                slices_[t].next.store(count * t / n_, std::memory_order_relaxed);
                slices_[t].end = count * (t + 1) / n_;
            }
            gen_++;
        }
        cv_.notify_all();
        work(0);
        std::unique_lock<std::mutex> lk(m_);
        done_.wait(lk, [this]() { return pending_ == 0; });
        fn_ = nullptr;
        if (!err_.empty()) throw std::runtime_error(err_);
    }

   private:
    struct alignas(64) Slice
    {
        std::atomic<size_t> next{0};
        size_t              end = 0;
    };
    void work(int self)
    {
        static const bool noAffinity = std::getenv("MPPB_POOL_NO_AFFINITY") != nullptr;  // A/B measurements only
        for (int v = 0; v < n_; v++)
        {
            Slice& sl = slices_[noAffinity ? v : (self + v) % n_];  // own slice first, then the others'
            for (;;)
            {
                const size_t i = sl.next.fetch_add(1, std::memory_order_relaxed);
                if (i >= sl.end) break;
                try
                {
                    (*fn_)(i);
                }
                catch (const std::exception& e)
                {
                    std::lock_guard<std::mutex> lk(m_);
                    if (err_.empty()) err_ = e.what();
                }
            }
        }
    }
    void loop(int self)
    {
        uint64_t seen = 0;
        for (;;)
        {
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&]() { return stop_ || gen_ != seen; });
                if (stop_) return;
                seen = gen_;
            }
            work(self);
            {
                std::lock_guard<std::mutex> lk(m_);
                pending_--;
            }
            done_.notify_one();
        }
    }
    int                                n_;
    std::vector<std::thread>           threads_;
    std::vector<Slice>                 slices_;
    std::mutex                         m_;
    std::condition_variable            cv_, done_;
    bool                               stop_ = false;
    uint64_t                           gen_  = 0;
    const std::function<void(size_t)>* fn_   = nullptr;
    int                                pending_ = 0;
    std::string                        err_;
};

// pinned host buffer (grow-only)
template <class T>
struct Pinned
{
    T*     p   = nullptr;
    size_t cap = 0;
    ~Pinned() { mppb_host_free(p); }
    T* reserve(size_t n)
    {
        if (n > cap)
        {
            mppb_host_free(p);
            p   = nullptr;
            cap = 0;
            void* q = nullptr;
            if (mppb_host_alloc(std::max<size_t>(n, 64) * sizeof(T) * 2, &q) != MPPB_OK)
                throw std::runtime_error("libmpp_b200: pinned host allocation failed");
            p   = static_cast<T*>(q);
            cap = std::max<size_t>(n, 64) * 2;
        }
        return p;
    }
};
}  // namespace

// Host self-test of the lattice index (no GPU): `ops` look-ups of random lattice coordinates (a small range, so that
// repeats are frequent, through several table growths); every coordinate must keep returning the node it was given
// the first time (stable address, same tag) and distinct coordinates distinct nodes. Returns the number of
// look-ups that did not (0 = correct).
int selftest_lattice_index(uint64_t seed, size_t ops, int range)
{
    std::pmr::monotonic_buffer_resource                       arena{1 << 16};
    LatticeIndex                                              grid{&arena};
    std::map<std::tuple<int32_t, int32_t, int32_t>, Node*>    ref;
    uint64_t                                                  x = seed * 0x9E3779B97F4A7C15ull + 1;
    auto                                                      rnd = [&x]() {
        x ^= x << 13;
        x ^= x >> 7;
        x ^= x << 17;
        return x;
    };
    const uint64_t r   = static_cast<uint64_t>(std::max(1, range));
    int            bad = 0;
    TNodeID        next = 0;
    for (size_t i = 0; i < ops; i++)
    {
        const int32_t cx = static_cast<int32_t>(rnd() % r) - static_cast<int32_t>(r / 2);
        const int32_t cy = static_cast<int32_t>(rnd() % r) - static_cast<int32_t>(r / 2);
        const int32_t cw = static_cast<int32_t>(rnd() % 16) - 8;
        Node&         n  = grid[NodeCoords(cx, cy, cw)];
        const auto    key = std::make_tuple(cx, cy, cw);
        const auto    it  = ref.find(key);
        if (it == ref.end())
        {
            if (n.id.has_value()) bad++;  // a fresh node must be untouched
            n.id     = next++;
            ref[key] = &n;
        }
        else if (it->second != &n)
            bad++;
    }
    // every node still holds its own tag (no two coordinates share a node)
    std::vector<char> seen(static_cast<size_t>(next), 0);
    for (const auto& kv : ref)
    {
        const Node* n = kv.second;
        if (!n->id.has_value() || *n->id >= next || seen[static_cast<size_t>(*n->id)]++) bad++;
    }
    return bad;
}

// Host self-test of CellOrder (no GPU): `trials` random sets of up to max_keys distinct lattice cells; returns the number
// of sets whose CellOrder iteration differs from the iteration order of the reference's
// std::unordered_map<NodeCoords, ..., NodeCoordsHash> fed the same keys in the same sequence (0 = identical).
int selftest_cell_order(uint64_t seed, size_t trials, int max_keys)
{
    uint64_t x   = seed * 0x9E3779B97F4A7C15ull + 1;
    auto     rnd = [&x]() {
        x ^= x << 13;
        x ^= x >> 7;
        x ^= x << 17;
        return x;
    };
    CellOrder      order;
    NodeCoordsHash hash;
    int            bad = 0;
    for (size_t t = 0; t < trials; t++)
    {
        const size_t                                              n = 1 + rnd() % static_cast<uint64_t>(std::max(1, max_keys));
        std::unordered_map<NodeCoords, uint32_t, NodeCoordsHash>  ref;
        std::vector<NodeCoords>                                   keys;
        // cells as an expansion reaches them: a small neighbourhood around some centre, occasionally far apart
        const int32_t cx = static_cast<int32_t>(rnd() % 2000) - 1000, cy = static_cast<int32_t>(rnd() % 2000) - 1000;
        const int32_t span = (rnd() % 4 == 0) ? 100000 : 24;
        order.reset();
        while (keys.size() < n)
        {
            const NodeCoords c(cx + static_cast<int32_t>(rnd() % static_cast<uint64_t>(span)) - span / 2,
                               cy + static_cast<int32_t>(rnd() % static_cast<uint64_t>(span)) - span / 2,
                               static_cast<int32_t>(rnd() % 15) - 7);
            if (ref.count(c)) continue;
            ref[c] = static_cast<uint32_t>(keys.size());
            keys.push_back(c);
            order.insert(hash(c));
        }
        uint16_t i  = order.first();
        bool     ok = true;
        for (const auto& kv : ref)
        {
            if (i == CellOrder::kNil || i != kv.second)
            {
                ok = false;
                break;
            }
            i = order.next(i);
        }
        if (i != CellOrder::kNil) ok = false;
        bad += ok ? 0 : 1;
    }
    return bad;
}

// Host self-test of the open set (no GPU): `ops` random insertions (keys drawn from `distinct_keys` values, so that
// equal keys are frequent) interleaved with pops; returns the number of pops that differ from a std::multimap
// driven the same way (0 = same order).
int selftest_open_set(uint64_t seed, size_t ops, int distinct_keys)
{
    std::vector<int>                    payload(ops);
    OpenSetT<int>                       heap;
    std::multimap<distance_t, int*>     ref;
    uint64_t                            x = seed * 0x9E3779B97F4A7C15ull + 1;
    auto                                rnd = [&x]() {
        x ^= x << 13;
        x ^= x >> 7;
        x ^= x << 17;
        return x;
    };
    int    bad = 0;
    size_t next = 0;
    while (next < ops || !ref.empty())
    {
        const bool doInsert = next < ops && (ref.empty() || rnd() % 3 != 0);
        if (doInsert)
        {
            const distance_t key = 0.25 * static_cast<double>(rnd() % static_cast<uint64_t>(std::max(1, distinct_keys)));
            payload[next]        = static_cast<int>(next);
            heap.insert(key, &payload[next]);
            ref.insert({key, &payload[next]});
            next++;
        }
        else
        {
            if (heap.empty() || heap.top() != ref.begin()->second) bad++;
            if (!heap.empty()) heap.pop();
            ref.erase(ref.begin());
        }
    }
    return bad + (heap.empty() ? 0 : 1);
}

// Host self-test of the worker pool (no GPU): `rounds` times, fn(i) for i in [0,count) on n_threads workers; returns the
// number of indices that were not executed exactly once per round (0 = correct), or -1 if an exception thrown by a
// work item is not rethrown to the caller.
int selftest_worker_pool(int n_threads, size_t count, int rounds)
{
    Workers                        pool(n_threads);
    std::vector<std::atomic<int>> hits(count);
    int                            bad = 0;
    for (int r = 0; r < rounds; r++)
    {
        for (auto& h : hits) h.store(0);
        pool.run(count, [&](size_t i) { hits[i].fetch_add(1); });
        for (auto& h : hits) bad += h.load() != 1;
    }
    if (count > 1)
    {
        bool thrown = false;
        try
        {
            pool.run(count, [&](size_t i) {
                if (i == count / 2) throw std::runtime_error("selftest");
            });
        }
        catch (const std::runtime_error&)
        {
            thrown = true;
        }
        if (!thrown) return -1;
        // the pool is usable after an exception
        for (auto& h : hits) h.store(0);
        pool.run(count, [&](size_t i) { hits[i].fetch_add(1); });
        for (auto& h : hits) bad += h.load() != 1;
    }
    return bad;
}

// ======================================================================================================
struct TPS_Astar::Impl
{
    TPS_Astar* self   = nullptr;
    mppb_ctx*  ctx    = nullptr;
    bool       ownCtx = false;
    const std::type_info* defaultHeuristicType = nullptr;  // target type of `heuristic` as the constructor set it

    // what is resident in the context. Each kind of resident data carries the context's epoch counter as it was
    // right after our upload: another planner on the same context, or a direct mppb_set_* call, bumps it.
    std::vector<std::pair<const ptg_t*, uint64_t>> uploadedPtgs;  // (object, ptg_t::revision())
    uint64_t                  ptgEpoch = 0;
    std::vector<int>          columnOffset;  // per PTG: first column of its collision-grid row, or -1
    // The reference re-reads ObstacleSource::obstacles() on every plan() (TPS_Astar.cpp:125-137). The cloud stays
    // resident here only while its CONTENT is unchanged: the key holds a hash of every coordinate (a cloud updated
    // in place — clear() plus the same number of insertPoint() calls, a freed buffer reallocated at the same
    // address — changes it), not buffer addresses.
    struct CloudKey
    {
        std::vector<size_t>   sizes;
        std::vector<uint64_t> hashes;
        bool                  hasBox = false;
        float                 box[4] = {0, 0, 0, 0};
        bool                  operator==(const CloudKey& o) const
        {
            return sizes == o.sizes && hashes == o.hashes && hasBox == o.hasBox && std::memcmp(box, o.box, sizeof(box)) == 0;
        }
    } uploadedCloud;
    bool     cloudValid = false;
    uint64_t cloudEpoch = 0;

    // evaluators resident in the context: object, its table and the parameters the device copy depends on.
    // A CostEvaluatorCostMap / PreferredWaypoint is immutable once built (as in the reference, whose grid and
    // waypoint map are private), so an unchanged key means the 4.8 MB costmap need not travel again.
    struct EvalKey
    {
        const void* obj  = nullptr;
        const void* data = nullptr;
        size_t      n    = 0;
        double      p[5] = {0, 0, 0, 0, 0};
        bool        operator==(const EvalKey& o) const
        {
            return obj == o.obj && data == o.data && n == o.n && std::memcmp(p, o.p, sizeof(p)) == 0;
        }
    };
    std::vector<EvalKey> uploadedEvals;
    bool                 evalsValid = false;
    uint64_t             evalEpoch  = 0;  // the context's evaluator epoch right after our upload

    // evaluator slots of the current plan
    std::vector<int> evalSlot;  // per costEvaluators_ entry: GPU slot or -1 (host plugin)
    int              numGpuEvals = 0;

    // device-side expansion (mppb_expand_nodes): available when every PTG is of the collision-grid kind and every cost
    // evaluator lives in a GPU slot; then phases A (second half), C, D and E leave the host altogether
    bool                     useExpand = false;
    uint32_t                 expandStride = 0;  // neighbour records per node in hNeighbors
    void                     setupExpand();

    Pinned<double>         hPoses, hHoloOut, hFrom, hRel, hCost;
    Pinned<float>          hRows, hBoxes;
    Pinned<mppb_holo_cand> hCands;
    Pinned<uint32_t>       hOffsets;

    struct Query;
    void                       ensureContext();
    void                       uploadPtgs(const TrajectoriesAndRobotShape& trs);
    void                       uploadCloud(const std::vector<ObstacleSource::Ptr>& srcs, const float* box);
    void                       uploadEvaluators();
    std::vector<PlannerOutput> run(const std::vector<const PlannerInput*>& ins, bool sharedCloud, int nThreads, bool materializeTrees);
};

// ---- one planning query: all the A* state of the reference's plan() ----------------------------------
struct TPS_Astar::Impl::Query
{
    TPS_Astar*                   planner = nullptr;
    const TPS_Astar_Parameters*  P       = nullptr;
    const PlannerInput*          in      = nullptr;
    PlannerOutput                po;
    TrajectoriesAndRobotShape    trs;  // the PTG objects this query mutates
    const std::vector<int>*      columnOffset = nullptr;
    Clock::time_point            t0;
    double                       tLastCallback = 0;

    // lattice and open set live in a per-query arena: node addresses stay stable (node-based containers),
    // allocation is a pointer bump and the whole search state is dropped at once when the query ends
    std::pmr::monotonic_buffer_resource                         searchArena{1 << 16};
    LatticeIndex                                                grid{&searchArena};
    OpenSetT<Node>                                              openSet;  // see OpenSetT
    TNodeID                                                nextFreeId = 0;
    Node*                                                  nodeGoal   = nullptr;
    NodeCoords                                             goalCellIndices;
    std::unordered_map<NodeCoords, normalized_speed_t, NodeCoordsHash> nodesWithDesiredSpeed;
    double                                                 MAX_XY_DIST = 0;
    Node*                                                  current     = nullptr;
    bool                                                   done        = false;
    bool                                                   active      = false;  // expands a node this round
    bool                                                   needF       = false;  // tentative edges wait for relaxation
    size_t                                                 costBase    = 0;      // first row of this query in the cost batch
    size_t                                                 slot        = 0;      // position in the round's node batch
    float                                                  box[4]      = {0, 0, 0, 0};

    // Tree insertions of phase F do not influence the search (open set, g/f scores and node states do): a single
    // query queues them and applies them, in order, while the GPU evaluates the next expansion's collisions.
    struct TreeOp
    {
        bool            rewire = false;
        TNodeID         parent = 0, child = 0;
        SE2_KinState    state;
        MoveEdgeSE2_TPS edge;
    };
    std::vector<TreeOp> treeOps;
    bool                deferTreeOps = false;
    bool                directHeuristic = false;  // planner->heuristic is still the constructor's default: call it directly
    cost_t heuristicOf(const SE2_KinState& st) const
    {
        return directHeuristic ? planner->default_heuristic(st, in->stateGoal) : planner->heuristic(st, in->stateGoal);
    }
    void flushTreeOps()
    {
        if (treeOps.empty()) return;
        auto&                                  tree        = po.motionTree;
        MotionPrimitivesTreeSE2::TListEdges*   parentEdges = nullptr;
        const MotionPrimitivesTreeSE2::node_t* parentNode  = nullptr;
        TNodeID                                curParent   = 0;
        for (TreeOp& op : treeOps)
        {
            if (op.rewire)
            {
                tree.rewire_node_parent(op.child, op.edge);
                continue;
            }
            // the children of one expansion share their parent: its edge list and tree node are looked up once
            if (!parentEdges || op.parent != curParent)
            {
                parentEdges = &tree.edges_to_children[op.parent];
                parentNode  = &tree.nodes().at(op.parent);
                curParent   = op.parent;
            }
            tree.insert_child_of(*parentEdges, *parentNode, op.parent, op.child, op.state, std::move(op.edge));
        }
        treeOps.clear();
    }

    // ---- device-side expansion path ----
    // The motion tree does not influence the search, so this path does not touch it while searching: what the tree
    // needs is kept in the lattice nodes themselves (see Node), and the tree — a map node plus a list node with a
    // 600-byte edge and its interpolated samples per tree node — is built once from them: when the query ends, or, for
    // a batch planned with materializeTrees == false, when somebody asks for it.
    size_t           nearTies  = 0;       // see GpuStats::holoNearTies
    bool             useExpand = false;
    bool             lazyTree  = false;   // leave po.motionTree empty; the caller builds it on demand
    CellOrder        cellOrder;           // see CellOrder
    mppb_expand_node expIn{};             // this round's record for mppb_expand_nodes
    uint32_t         relaxSeq     = 0;    // accepted relaxations so far
    size_t           treeNodes    = 1;    // nodes the tree would hold now (the root + every node relaxed at least once)
    size_t           treeParents  = 0;    // nodes that own an edge list
    const Node*      bestNode     = nullptr;
    TNodeID          rootId       = 0;
    size_t           nSeg         = 0;    // pathInterpolatedSegments (the planner's parameters may be gone later)
    std::optional<std::pair<TNodeID, TPose2D>> goalSnap;  // node whose pose popNext() moved onto the exact goal

    // phase scratch
    std::pmr::monotonic_buffer_resource arena{16384};
    std::vector<Candidate>        cands;
    std::vector<TNavDynamicState> dynOfPtg;
    std::vector<mppb_holo_cand>   holoCands;
    std::vector<PendingEdge>      pending;

    // TPS_Astar.h:224-230: the lattice index takes its argument as float ([MRPT] wrapToPi<float>)
    int32_t x2idx(float x) const { return static_cast<int32_t>(x / P->grid_resolution_xy); }
    int32_t phi2idx(float yaw) const
    {
        const float pi = static_cast<float>(M_PI), twoPi = static_cast<float>(2.0 * M_PI);
        float       a   = yaw + pi;
        const bool  neg = a < 0;
        a               = std::fmod(a, twoPi);
        if (neg) a += twoPi;
        return static_cast<int32_t>((a - pi) / P->grid_resolution_yaw);
    }
    NodeCoords nodeGridCoords(const TPose2D& p) const { return NodeCoords(x2idx(p.x), x2idx(p.y), phi2idx(p.phi)); }
    NodeCoords nodeGridCoords(const TPoint2D& p) const { return NodeCoords(x2idx(p.x), x2idx(p.y)); }

    Node& getOrCreateNodeByPose(const SE2_KinState& p)
    {
        Node& n = grid[nodeGridCoords(p.pose)];
        if (!n.id.has_value())
        {
            n.id    = nextFreeId++;
            n.state = p;
        }
        return n;
    }

    void start()
    {
        t0 = Clock::now();
        MPP_ASSERT(in->ptgs.initialized());
        MPP_ASSERT(in->worldBboxMin != in->worldBboxMax);
        MPP_ASSERT(within_bbox(in->stateStart.pose, in->worldBboxMax, in->worldBboxMin));
        MPP_ASSERT(!in->stateGoal.state.isEmpty());
        if (in->stateGoal.state.isPoint())
            MPP_ASSERT(within_bbox(in->stateGoal.state.point(), in->worldBboxMax, in->worldBboxMin));
        else
            MPP_ASSERT(within_bbox(in->stateGoal.state.pose(), in->worldBboxMax, in->worldBboxMin));
        po.originalInput = *in;
        for (const auto& ptg : trs.ptgs) MAX_XY_DIST = std::max(MAX_XY_DIST, ptg->getRefDistance());
        MPP_ASSERT(MAX_XY_DIST > 0);
        // world box as apply_clipping_box sees it: double -> float (ObstacleSource.cpp:29-30)
        box[0] = static_cast<float>(in->worldBboxMin.x);
        box[1] = static_cast<float>(in->worldBboxMin.y);
        box[2] = static_cast<float>(in->worldBboxMax.x);
        box[3] = static_cast<float>(in->worldBboxMax.y);

        auto& tree = po.motionTree;
        {
            Node& n   = getOrCreateNodeByPose(in->stateStart);
            n.state   = in->stateStart;
            tree.root = n.id.value();
            rootId    = tree.root;
            nSeg      = P->pathInterpolatedSegments;
            if (!useExpand) tree.insert_root_node(tree.root, n.state);
            n.gScore           = 0;
            n.fScore           = heuristicOf(n.state);
            n.pendingInOpenSet = true;
            openSet.insert(n.fScore, &n);
        }
        nodeGoal      = &getOrCreateNodeByPose(in->stateGoal.asSE2KinState());
        po.goalNodeId = nodeGoal->id.value();
        goalCellIndices = in->stateGoal.state.isPoint() ? nodeGridCoords(in->stateGoal.state.point())
                                                        : nodeGridCoords(in->stateGoal.state.pose());
        nodesWithDesiredSpeed[goalCellIndices] = 0;  // goal speed = 0
    }

    // head of the A* iteration (TPS_Astar.cpp:193-255): pick the open node with the lowest fScore,
    // finish if it lies in the goal cell. Returns false when the query has ended.
    bool popNext()
    {
        if (done) return false;
        if (openSet.empty())
        {
            finish();
            return false;
        }
        Node& cur = *openSet.top();
        if (nodeGridCoords(cur.state.pose).sameLocation(goalCellIndices))
        {
            flushTreeOps();
            // snap the node onto the exact goal; refine_trajectory() fixes the last edge later
            TPose2D snapped = cur.state.pose;
            if (in->stateGoal.state.isPoint())
            {
                const TPoint2D g        = in->stateGoal.state.point();
                snapped.x               = g.x;
                snapped.y               = g.y;
                nodeGoal                = &cur;
                po.goalNodeId           = cur.id.value();
                po.bestNodeId           = po.goalNodeId;
                bestNode                = &cur;
                po.bestNodeIdCostToGoal = 0;
            }
            else
            {
                snapped                 = in->stateGoal.state.pose();
                po.bestNodeIdCostToGoal = 0;
            }
            if (useExpand)
                goalSnap = std::make_pair(*cur.id, snapped);  // the tree node moves; the edge keeps its stateTo (= cur.state)
            else
            {
                cur.state.pose                         = snapped;
                po.motionTree.node_state(*cur.id).pose = cur.state.pose;
            }
            finish();
            return false;
        }
        cur.pendingInOpenSet = false;
        cur.visited          = true;
        openSet.pop();
        current = &cur;
        po.nodesExpanded++;
        return true;
    }

    void finish()
    {
        flushTreeOps();
        done       = true;
        current    = nullptr;
        po.success = po.goalNodeId == po.bestNodeId;
        if (useExpand)
        {
            if (bestNode) po.pathCost = bestNode->gScore;  // a node's tree cost is its gScore (see Node)
            po.treeNodes   = treeNodes;
            po.treeParents = treeParents;
            po.pathLength  = 0;
            for (const Node* n = bestNode; n; n = n->cameFrom) po.pathLength++;
            if (!lazyTree) buildTree(po.motionTree, *in);
        }
        else
        {
            if (po.bestNodeId) po.pathCost = po.motionTree.nodes().at(*po.bestNodeId).cost_;
            po.treeNodes   = po.motionTree.nodes().size();
            po.treeParents = po.motionTree.edges_to_children.size();
            po.pathLength  = 0;
            if (po.bestNodeId)
                for (std::optional<TNodeID> id = po.bestNodeId; id; id = po.motionTree.nodes().at(*id).parentID_) po.pathLength++;
        }
        po.computationTime = seconds_since(t0);
    }

    // ---- phase A: candidate TPS points of `current` (TPS_Astar.cpp:549-736) ----
    // In two parts, so that the GPU can work while the host finishes: phaseAPrepare() lists the TPS points of every
    // PTG and collects what the collision kernels need (the HolonomicBlend candidate records; the collision-grid
    // kernel needs the node pose only); phaseACandidates() then reconstructs the end poses and lattice cells. The
    // PTG objects see the same sequence of state changes as in the reference's single loop.
    struct TpsPoint
    {
        trajectory_index_t k;
        ptg_step_t         step;
        normalized_speed_t speed;
        uint32_t           evalIndex;  // HolonomicBlend: index into holoCands
    };
    struct TpsRange  // the TPS points of one PTG inside tpsAll
    {
        size_t             begin = 0, end = 0, nTargets = 0;
        bool               trimmable = false, dedup = false;
        uint32_t           copies     = 1;
        normalized_speed_t firstSpeed = 1.0;
    };
    std::vector<TpsPoint> tpsAll;
    std::vector<TpsRange> tpsOfPtg;

    void phaseAPrepare()
    {
        cands.clear();
        holoCands.clear();
        tpsAll.clear();
        tpsOfPtg.assign(trs.ptgs.size(), TpsRange());
        dynOfPtg.assign(trs.ptgs.size(), TNavDynamicState());
        const Node&      from        = *current;
        const NodeCoords iGoalCoords = goalCellIndices;
        const TPose2D    relGoal     = in->stateGoal.asSE2KinState().pose - from.state.pose;

        arena.release();  // phase scratch of the previous expansion
        std::pmr::map<std::pair<trajectory_index_t, double>, uint32_t> holoIndex(&arena);

        for (size_t ptgIdx = 0; ptgIdx < trs.ptgs.size(); ptgIdx++)
        {
            ptg_t& ptg = *trs.ptgs[ptgIdx];
            MPP_ASSERT(ptg.isInitialized());
            const bool               trimmable = ptg.isSpeedTrimmable();
            const duration_seconds_t ptg_dt    = ptg.getPathStepDuration();
            {
                TNavDynamicState ds;
                (ds.curVelLocal = from.state.vel).rotate(-from.state.pose.phi);
                ds.relTarget        = relGoal;
                const auto it       = nodesWithDesiredSpeed.find(iGoalCoords);
                ds.targetRelSpeed   = it != nodesWithDesiredSpeed.end() ? it->second : 0;
                ptg.updateNavDynamicState(ds);
            }
            dynOfPtg[ptgIdx] = ptg.getCurrentNavDynamicState();

            // std::set<trajectory_index_t> of the reference as a sorted vector without duplicates (same iteration order);
            // the direct-to-goal TPS points are the first nTargets entries of the PTG's range
            std::pmr::vector<trajectory_index_t> trajIdxs(&arena);
            auto insertTraj = [&trajIdxs](trajectory_index_t k) {
                const auto it = std::lower_bound(trajIdxs.begin(), trajIdxs.end(), k);
                if (it == trajIdxs.end() || *it != k) trajIdxs.insert(it, k);
            };
            TpsRange& R = tpsOfPtg[ptgIdx];
            R.begin     = tpsAll.size();
            R.trimmable = trimmable;
            MPP_ASSERT(P->max_ptg_trajectories_to_explore >= 2);
            for (size_t i = 0; i < P->max_ptg_trajectories_to_explore; i++)
            {
                // integer division first, then round (TPS_Astar.cpp:619-623)
                const size_t q = i * (ptg.getPathCount() - 1) / (P->max_ptg_trajectories_to_explore - 1);
                insertTraj(round_i(static_cast<double>(q)));
            }
            std::pmr::vector<normalized_speed_t> speeds(&arena);
            if (trimmable)
            {
                MPP_ASSERT(P->max_ptg_speeds_to_explore >= 1);
                const normalized_speed_t inc = 1.0 / P->max_ptg_speeds_to_explore;
                for (normalized_speed_t s = inc; s < 1.001; s += inc) speeds.push_back(s);
            }
            else
                speeds.push_back(1.0);
            {
                // the trajectory straight to the goal, when the PTG can reach it
                int    goalK = 0;
                double goalD = 0;
                if (ptg.inverseMap_WS2TP(relGoal.x, relGoal.y, goalK, goalD, P->grid_resolution_xy))
                {
                    ptg_step_t goalStep = 0;
                    // the reference passes the normalised distance here (TPS_Astar.cpp:658)
                    if (ptg.getPathStepForDist(static_cast<uint16_t>(goalK), goalD, goalStep))
                        for (const auto s : speeds)
                        {
                            tpsAll.push_back({goalK, goalStep, s, 0u});
                            R.nTargets = tpsAll.size() - R.begin;
                        }
                    insertTraj(goalK);
                }
            }
            for (const auto s : speeds)
                for (const auto k : trajIdxs)
                    for (const duration_seconds_t t : P->ptg_sample_timestamps)
                    {
                        const ptg_step_t step     = static_cast<ptg_step_t>(round_i(t / ptg_dt));
                        const size_t     maxSteps = ptg.getPathStepCount(static_cast<uint16_t>(k));
                        MPP_ASSERT(maxSteps >= 1);
                        if (step >= maxSteps) continue;
                        tpsAll.push_back({k, step, s, 0u});
                    }
            R.end = tpsAll.size();

            // A trimmable PTG whose paths ignore the trimmed speed yields, for every speed after the first, exact copies
            // of the first speed's TPS points (same k, step, distance, pose, cell). Copies can never win the strict "<"
            // of the per-cell selection nor change the goal-cell set, so only the first speed is carried on; `copies`
            // keeps the reference's count of evaluated points.
            R.dedup      = trimmable && !ptg.pathsDependOnTrimmableSpeed() && speeds.size() > 1;
            R.copies     = R.dedup ? static_cast<uint32_t>(speeds.size()) : 1u;
            R.firstSpeed = speeds.front();

            if (ptg.kind() == PtgKind::CollisionGrid) continue;
            // HolonomicBlend: one candidate record per (k, trimmed speed) of the PTG's TPS points. (Records are made
            // before the world-box test of phaseACandidates(): a point that fails it leaves an unused record.)
            holoIndex.clear();
            for (size_t i = R.begin; i < R.end; i++)
            {
                TpsPoint& tp = tpsAll[i];
                if (tp.step == 0) continue;  // no motion
                if (trimmable && tp.speed != ptg.trimmableSpeed_) ptg.trimmableSpeed_ = tp.speed;
                const auto key = std::make_pair(tp.k, tp.speed);
                auto       it  = holoIndex.find(key);
                if (it == holoIndex.end())
                {
                    mppb_holo_cand hc;
                    static_cast<const HolonomicBlend&>(ptg).holoParams(static_cast<uint16_t>(tp.k), hc);
                    hc.node      = 0;  // position in the round's node batch: filled in when the batch is assembled
                    hc.ptg_index = static_cast<int32_t>(ptgIdx);
                    holoCands.push_back(hc);
                    it = holoIndex.emplace(key, static_cast<uint32_t>(holoCands.size() - 1)).first;
                }
                tp.evalIndex = it->second;
            }
        }
    }

    void phaseACandidates()
    {
        const Node&    from     = *current;
        const double   halfCell = P->grid_resolution_xy * 0.5;
        const TPose2D& lo = in->worldBboxMin, &hi = in->worldBboxMax;
        // from.state.pose (+) relPose with the heading's cos/sin taken once (TPose2D::operator+ takes them per call)
        const double fromC = std::cos(from.state.pose.phi), fromS = std::sin(from.state.pose.phi);
        for (size_t ptgIdx = 0; ptgIdx < trs.ptgs.size(); ptgIdx++)
        {
            ptg_t&          ptg       = *trs.ptgs[ptgIdx];
            const TpsRange& R         = tpsOfPtg[ptgIdx];
            const bool      trimmable = R.trimmable, dedup = R.dedup;
            const bool      onGrid    = ptg.kind() == PtgKind::CollisionGrid;
            // initTPObstacleSingle(k) depends on (k, trimmed speed, PTG state) only: consecutive TPS points of the same
            // path and speed (the sample timestamps) share it
            trajectory_index_t initK     = -1;
            normalized_speed_t initSpeed = -1;
            distance_t         initVal   = 0;
            for (size_t i = R.begin; i < R.end; i++)
            {
                const TpsPoint& tp = tpsAll[i];
                if (tp.step == 0) continue;  // no motion
                if (trimmable && tp.speed != ptg.trimmableSpeed_) ptg.trimmableSpeed_ = tp.speed;
                if (dedup && tp.speed != R.firstSpeed) continue;
                const uint16_t   k16        = static_cast<uint16_t>(tp.k);
                const distance_t relTrgDist = ptg.getPathDist(k16, tp.step);
                const TPose2D    relPose    = ptg.getPathPose(k16, tp.step);
                const TPose2D    absPose(from.state.pose.x + relPose.x * fromC - relPose.y * fromS,
                                         from.state.pose.y + relPose.x * fromS + relPose.y * fromC,
                                         wrap_to_pi(from.state.pose.phi + relPose.phi));
                if (absPose.x < lo.x || absPose.y < lo.y || absPose.phi < lo.phi) continue;
                if (absPose.x > hi.x - halfCell || absPose.y > hi.y - halfCell || absPose.phi > hi.phi) continue;
                Candidate c;
                c.ptgIdx     = ptgIdx;
                c.k          = tp.k;
                c.step       = tp.step;
                c.relTrgDist = relTrgDist;
                c.relPose    = relPose;
                c.nc         = nodeGridCoords(absPose);
                if (tp.k != initK || tp.speed != initSpeed)
                {
                    initVal   = ptg.initTPObstacleSingle(k16);
                    initK     = tp.k;
                    initSpeed = tp.speed;
                }
                c.initDist   = initVal;
                c.isTarget   = i - R.begin < R.nTargets;
                c.onGrid     = onGrid;
                c.copies     = R.copies;
                c.evalIndex  = onGrid ? static_cast<uint32_t>((*columnOffset)[ptgIdx] + tp.k) : tp.evalIndex;
                cands.push_back(c);
            }
        }
    }

    // ---- phases C + D: neighbour selection and tentative edges (TPS_Astar.cpp:738-814, 269-341) ----
    void phaseCD(const float* gridRow, uint32_t gridCols, const double* holoOut)
    {
        // same containers and hash as the reference (their iteration order decides node ids), fed from a
        // per-expansion arena: the allocator does not influence bucket layout or iteration order
        arena.release();
        std::pmr::unordered_map<NodeCoords, PathToNeighbor, NodeCoordsHash> bestPaths(&arena);
        std::pmr::unordered_set<NodeCoords, NodeCoordsHash>                 goalNodeCoords(&arena);
        size_t                                                         lastPtg = static_cast<size_t>(-1);
        for (const Candidate& c : cands)
        {
            if (c.ptgIdx != lastPtg)
            {
                goalNodeCoords.clear();  // the reference declares this set inside the per-PTG loop
                lastPtg = c.ptgIdx;
            }
            po.tpsPointsEvaluated += c.copies;
            MPP_ASSERT(!c.onGrid || c.evalIndex < gridCols);
            const double     raw          = c.onGrid ? static_cast<double>(gridRow[c.evalIndex]) : holoOut[c.evalIndex];
            const distance_t freeDistance = std::min(c.initDist, raw);
            // decision guard (HolonomicBlend distances come out of a quartic whose trigonometric branch uses the
            // device's acos/cos): count the decisions taken within 1e-9 relative of a tie
            if (!c.onGrid && raw < c.initDist && std::fabs(c.relTrgDist - freeDistance) < 1e-9 * freeDistance) nearTies++;
            if (c.relTrgDist >= freeDistance) continue;  // blocked
            if (c.isTarget)
                goalNodeCoords.insert(c.nc);
            else if (goalNodeCoords.count(c.nc) != 0)
                continue;  // a direct-to-goal path already owns this cell
            PathToNeighbor& path = bestPaths[c.nc];
            if (c.relTrgDist < path.ptgDist)
            {
                path.ptgDist         = c.relTrgDist;
                path.ptgIndex        = c.ptgIdx;
                path.ptgTrajIndex    = c.k;
                path.relReconstrPose = c.relPose;
                path.relTrgStep      = c.step;
                path.ptgDynState     = dynOfPtg[c.ptgIdx];
            }
        }

        pending.clear();
        const TPose2D& lo = in->worldBboxMin, &hi = in->worldBboxMax;
        // every neighbour is reached from `current`: the cos/sin of its heading (pose composition, twist rotation) and
        // the goal as seen from it are taken once, not once per neighbour (same libm calls on the same argument)
        const double  curC = std::cos(current->state.pose.phi), curS = std::sin(current->state.pose.phi);
        const TPose2D goalFromCurrent = in->stateGoal.asSE2KinState().pose - current->state.pose;
        for (const auto& kv : bestPaths)  // unordered_map order, as in the reference
        {
            const PathToNeighbor& nb = kv.second;
            if (!nb.ptgIndex.has_value()) continue;
            ptg_t& ptg = *trs.ptgs.at(*nb.ptgIndex);
            ptg.updateNavDynamicState(nb.ptgDynState);
            if (ptg.isSpeedTrimmable()) ptg.trimmableSpeed_ = nb.ptgTrimmableSpeed;
            const TPose2D  q_i      = current->state.pose.composeCS(curC, curS, nb.relReconstrPose);
            const TTwist2D relTwist = ptg.getPathTwist(static_cast<uint16_t>(nb.ptgTrajIndex), nb.relTrgStep);
            SE2_KinState   x_i;
            x_i.pose = q_i;
            (x_i.vel = relTwist).rotateCS(curC, curS);
            if (q_i.x < lo.x || q_i.y < lo.y || q_i.phi < lo.phi) continue;
            if (q_i.x > hi.x || q_i.y > hi.y || q_i.phi > hi.phi) continue;
            Node& neighborNode = getOrCreateNodeByPose(x_i);
            if (neighborNode.visited) continue;

            PendingEdge pe;
            pe.neighbor = &neighborNode;
            pe.x_i      = x_i;
            MoveEdgeSE2_TPS& e     = pe.edge;
            e.parentId             = current->id.value();
            e.ptgDist              = nb.ptgDist;
            e.ptgIndex             = static_cast<int8_t>(*nb.ptgIndex);
            e.ptgPathIndex         = static_cast<int16_t>(nb.ptgTrajIndex);
            e.ptgTrimmableSpeed    = nb.ptgTrimmableSpeed;
            e.ptgFinalGoalRelSpeed = 0;
            e.ptgFinalRelativeGoal = goalFromCurrent;
            e.stateFrom            = current->state;
            e.stateTo              = x_i;
            e.ptgStep              = nb.relTrgStep;
            edge_interpolated_path(e, trs, nb.relReconstrPose, nb.relTrgStep, P->pathInterpolatedSegments);
            pending.push_back(std::move(pe));
        }
    }

    // ---- phase F: relaxation (TPS_Astar.cpp:343-451) ----
    // gpuCosts: [pending.size()][numGpuEvals] evaluator values of this query's tentative edges
    void phaseF(const double* gpuCosts, int numGpuEvals, const std::vector<int>& evalSlot)
    {
        auto& tree = po.motionTree;
        // all tentative edges of this round leave `current`: its edge list and tree node are looked up once
        MotionPrimitivesTreeSE2::TListEdges*   parentEdges = nullptr;
        const MotionPrimitivesTreeSE2::node_t* parentNode  = nullptr;
        for (size_t i = 0; i < pending.size(); i++)
        {
            PendingEdge&     pe      = pending[i];
            MoveEdgeSE2_TPS& newEdge = pe.edge;
            Node&            nbNode  = *pe.neighbor;
            // Planner::cost_path_segment: execution time + evaluators in their registered order
            cost_t c = newEdge.estimatedExecTime;
            for (size_t e = 0; e < planner->costEvaluators_.size(); e++)
            {
                if (evalSlot[e] >= 0)
                    c += gpuCosts[i * numGpuEvals + evalSlot[e]];
                else
                    c += (*planner->costEvaluators_[e])(newEdge);  // user plugin on the host
            }
            newEdge.cost = c;
            po.edgesCosted++;
            MPP_ASSERT(newEdge.cost > .0);

            const cost_t tentative_gScore = current->gScore + newEdge.cost;
            if (tentative_gScore >= nbNode.gScore) continue;
            const bool hasToRewire = nbNode.cameFrom != nullptr;
            nbNode.cameFrom        = current;
            nbNode.gScore          = tentative_gScore;
            // NOTE: the heuristic is taken at the node's previous state, before it is overwritten
            const cost_t costToGoal = heuristicOf(nbNode.state);
            nbNode.fScore           = tentative_gScore + costToGoal;
            if (!nbNode.pendingInOpenSet)
            {
                nbNode.pendingInOpenSet = true;
                openSet.insert(nbNode.fScore, &nbNode);
            }
            nbNode.state = pe.x_i;
            if (deferTreeOps)
            {
                TreeOp op;
                op.rewire = hasToRewire;
                op.parent = newEdge.parentId;
                op.child  = nbNode.id.value();
                op.state  = nbNode.state;
                op.edge   = std::move(newEdge);  // `pending` is cleared below
                treeOps.push_back(std::move(op));
            }
            else if (hasToRewire)
                tree.rewire_node_parent(nbNode.id.value(), newEdge);
            else
            {
                if (!parentEdges)
                {
                    parentEdges = &tree.edges_to_children[newEdge.parentId];
                    parentNode  = &tree.nodes().at(newEdge.parentId);
                }
                // (the edge moves into the tree: `pending` is cleared below)
                tree.insert_child_of(*parentEdges, *parentNode, newEdge.parentId, nbNode.id.value(), nbNode.state,
                                     std::move(newEdge));
            }
            if (costToGoal < po.bestNodeIdCostToGoal)
            {
                po.bestNodeIdCostToGoal = costToGoal;
                po.bestNodeId           = nbNode.id.value();
            }
        }
        pending.clear();

        const double tNow = seconds_since(t0);
        if (tNow > P->maximumComputationTime || po.nodesExpanded >= P->maxExpansions)
        {
            finish();  // timeout: the best node so far is the result
            return;
        }
        if (planner->progressCallback_ && (tNow - tLastCallback) > planner->progressCallbackCallPeriod_ && po.bestNodeId)
        {
            tLastCallback = tNow;
            flushTreeOps();
            auto [foundPath, pathEdges] = tree.backtrack_path(*po.bestNodeId);
            ProgressCallbackData pcd;
            pcd.bestCostFromStart = tree.nodes().at(*po.bestNodeId).cost_;
            pcd.bestCostToGoal    = po.bestNodeIdCostToGoal;
            pcd.bestFinalNode     = po.bestNodeId;
            pcd.bestPath          = std::move(pathEdges);
            pcd.costEvaluators    = &planner->costEvaluators_;
            pcd.originalPlanInput = in;
            pcd.tree              = &tree;
            planner->progressCallback_(pcd);
        }
    }

    // ======== device-side expansion path ========
    // ---- phase A, what is left of it: PTG state, and the path straight to the goal per PTG (TPS_Astar.cpp:587-611,
    // :650-677; inverseMap_WS2TP takes atan2 from the host C library) ----
    void phaseAExpand()
    {
        const Node& from = *current;
        // cos/sin of the node's heading, once: for the goal as seen from the node, for the local velocity, and for the
        // expansion call (which would otherwise take them for every node of the batch on one thread)
        const double  fromC = std::cos(from.state.pose.phi), fromS = std::sin(from.state.pose.phi);
        const TPose2D relGoal = in->stateGoal.asSE2KinState().pose.minusCS(from.state.pose, fromC, fromS);
        expIn.cos_phi = fromC;
        expIn.sin_phi = fromS;
        for (int i = 0; i < 3; i++)
        {
            const double* lo = &in->worldBboxMin.x;  // (x, y, phi are consecutive doubles of TPose2D)
            const double* hi = &in->worldBboxMax.x;
            expIn.bbox_min[i] = lo[i];
            expIn.bbox_max[i] = hi[i];
        }
        expIn.pose[0] = from.state.pose.x;
        expIn.pose[1] = from.state.pose.y;
        expIn.pose[2] = from.state.pose.phi;
        const auto itSpeed = nodesWithDesiredSpeed.find(goalCellIndices);
        for (size_t ptgIdx = 0; ptgIdx < trs.ptgs.size(); ptgIdx++)
        {
            ptg_t& ptg = *trs.ptgs[ptgIdx];
            {
                TNavDynamicState ds;
                (ds.curVelLocal = from.state.vel).rotateCS(fromC, -fromS);  // rotate(-phi): cos is even, sin is odd
                ds.relTarget      = relGoal;
                ds.targetRelSpeed = itSpeed != nodesWithDesiredSpeed.end() ? itSpeed->second : 0;
                ptg.updateNavDynamicState(ds);
            }
            if (ptg.isSpeedTrimmable()) ptg.trimmableSpeed_ = 1.0;  // where the reference's loops leave it (:721-722, :282-284)
            int    goalK = 0;
            double goalD = 0;
            expIn.goal_k[ptgIdx]    = -1;
            expIn.goal_step[ptgIdx] = 0xffffffffu;
            if (ptg.inverseMap_WS2TP(relGoal.x, relGoal.y, goalK, goalD, P->grid_resolution_xy))
            {
                expIn.goal_k[ptgIdx] = goalK;
                ptg_step_t goalStep  = 0;
                // the reference passes the normalised distance here (TPS_Astar.cpp:658)
                if (ptg.getPathStepForDist(static_cast<uint16_t>(goalK), goalD, goalStep)) expIn.goal_step[ptgIdx] = goalStep;
            }
        }
    }

    // ---- what phases C, D and F leave to the host (TPS_Astar.cpp:783-814, :303-314, :343-451): the reference's hash
    // map decides the ORDER in which the reached cells become nodes (its iteration order assigns node ids), then the
    // visited test and the relaxation of each neighbour. recs: the node's reached cells in first-insertion order. ----
    void consumeNeighbors(const mppb_neighbor* recs, uint32_t n, uint32_t tpsPoints)
    {
        po.tpsPointsEvaluated += tpsPoints;
        arena.release();
        // The neighbours' lattice nodes are scattered over a search state far larger than the caches: three passes, so
        // that the misses overlap instead of queueing up behind each other — (1) the reference's map of reached cells
        // (its iteration order is the order in which cells become nodes) while the index slots are prefetched,
        // (2) the existing nodes are found and their first cache line is prefetched, (3) the ordered pass.
        cellOrder.reset();
        {
            MPPB_PROF(2);
            for (uint32_t i = 0; i < n; i++)
            {
                const mppb_neighbor& r = recs[i];
                cellOrder.insert(NodeCoordsHash()(NodeCoords(r.cell[0], r.cell[1], r.cell[2])));
                grid.prefetch(r.cell[0], r.cell[1], r.cell[2]);
            }
        }
        Node** known = static_cast<Node**>(arena.allocate(sizeof(Node*) * std::max<uint32_t>(n, 1), alignof(Node*)));
        {
            MPPB_PROF(3);
            for (uint32_t i = 0; i < n; i++)
            {
                const mppb_neighbor& r = recs[i];
                known[i]               = grid.find(r.cell[0], r.cell[1], r.cell[2]);
                if (known[i])
                {
                    __builtin_prefetch(known[i]);
                    if (kPrefetchStateLine) __builtin_prefetch(reinterpret_cast<const char*>(known[i]) + 64, 1);
                }
            }
        }
        MPPB_PROF(4);
        for (uint16_t ri = cellOrder.first(); ri != CellOrder::kNil; ri = cellOrder.next(ri))  // the reference's map order
        {
            const mppb_neighbor& r = recs[ri];
            if (!r.in_bbox) continue;
            Node& nbNode = known[ri] ? *known[ri] : grid[NodeCoords(r.cell[0], r.cell[1], r.cell[2])];
            if (!nbNode.id.has_value())
            {
                nbNode.id         = nextFreeId++;
                nbNode.state.pose = TPose2D(r.pose[0], r.pose[1], r.pose[2]);
                nbNode.state.vel  = TTwist2D(r.vel[0], r.vel[1], r.vel[2]);
            }
            if (nbNode.visited) continue;
            const cost_t edgeCost = r.cost;
            po.edgesCosted++;
            MPP_ASSERT(edgeCost > .0);
            const cost_t tentative_gScore = current->gScore + edgeCost;
            if (tentative_gScore >= nbNode.gScore) continue;
            const bool firstTime = nbNode.cameFrom == nullptr;
            nbNode.cameFrom      = current;
            nbNode.gScore        = tentative_gScore;
            // NOTE: the heuristic is taken at the node's previous state, before it is overwritten
            cost_t costToGoal;
            {
                MPPB_PROF(5);
                costToGoal = heuristicOf(nbNode.state);
            }
            nbNode.fScore = tentative_gScore + costToGoal;
            if (!nbNode.pendingInOpenSet)
            {
                MPPB_PROF(6);
                nbNode.pendingInOpenSet = true;
                openSet.insert(nbNode.fScore, &nbNode);
            }
            nbNode.state.pose = TPose2D(r.pose[0], r.pose[1], r.pose[2]);
            nbNode.state.vel  = TTwist2D(r.vel[0], r.vel[1], r.vel[2]);
            // the node's place in the tree (insert_node_and_edge / rewire_node_parent of the reference, :380-398)
            nbNode.edgePtg  = r.ptg_index;
            nbNode.edgeK    = r.k;
            nbNode.edgeStep = r.step;
            nbNode.edgeDist = r.ptg_dist;
            nbNode.edgeCost = edgeCost;
            nbNode.edgeSeq  = relaxSeq++;
            if (firstTime)
            {
                nbNode.firstState = nbNode.state;
                treeNodes++;
            }
            if (!current->everParent)
            {
                current->everParent = true;
                treeParents++;
            }
            if (costToGoal < po.bestNodeIdCostToGoal)
            {
                po.bestNodeIdCostToGoal = costToGoal;
                po.bestNodeId           = nbNode.id.value();
                bestNode                = &nbNode;
            }
        }
        const double tNow = seconds_since(t0);
        if (tNow > P->maximumComputationTime || po.nodesExpanded >= P->maxExpansions)
        {
            finish();  // timeout: the best node so far is the result
            return;
        }
        if (planner->progressCallback_ && (tNow - tLastCallback) > planner->progressCallbackCallPeriod_ && po.bestNodeId)
        {
            tLastCallback = tNow;
            MotionPrimitivesTreeSE2 tree;  // the tree as it stands now
            buildTree(tree, *in);
            auto [foundPath, pathEdges] = tree.backtrack_path(*po.bestNodeId);
            ProgressCallbackData pcd;
            pcd.bestCostFromStart = tree.nodes().at(*po.bestNodeId).cost_;
            pcd.bestCostToGoal    = po.bestNodeIdCostToGoal;
            pcd.bestFinalNode     = po.bestNodeId;
            pcd.bestPath          = std::move(pathEdges);
            pcd.costEvaluators    = &planner->costEvaluators_;
            pcd.originalPlanInput = in;
            pcd.tree              = &tree;
            planner->progressCallback_(pcd);
        }
    }

    // The motion tree of the reference from the lattice nodes (see Node): every relaxed node with its last accepted edge,
    // in the order of the accepting relaxations.
    void buildTree(MotionPrimitivesTreeSE2& tree, const PlannerInput& input) const
    {
        tree      = MotionPrimitivesTreeSE2();
        tree.root = rootId;
        tree.insert_root_node(rootId, input.stateStart);
        std::vector<const Node*> order;
        order.reserve(treeNodes);
        grid.forEachNode([&](const Node& n) {
            if (n.cameFrom) order.push_back(&n);
            // (a parent whose children were all re-parented later still owns an — empty — edge list, as after the
            // reference's rewire_node_parent)
            if (n.everParent) tree.edges_to_children[n.id.value()];
        });
        std::sort(order.begin(), order.end(), [](const Node* a, const Node* b) { return a->edgeSeq < b->edgeSeq; });
        const TPose2D                          goalPose    = input.stateGoal.asSE2KinState().pose;
        const Node*                            curParent   = nullptr;
        MotionPrimitivesTreeSE2::TListEdges*   parentEdges = nullptr;
        const MotionPrimitivesTreeSE2::node_t* parentNode  = nullptr;
        TPose2D                                goalRel;
        for (const Node* n : order)
        {
            const Node&   par      = *n->cameFrom;
            const TNodeID parentId = par.id.value();
            if (&par != curParent)
            {
                parentEdges = &tree.edges_to_children[parentId];
                parentNode  = &tree.nodes().at(parentId);
                goalRel     = goalPose - par.state.pose;  // the goal as seen from the parent when it was expanded
                curParent   = &par;
            }
            MoveEdgeSE2_TPS e;
            e.parentId             = parentId;
            e.ptgDist              = n->edgeDist;
            e.ptgIndex             = static_cast<int8_t>(n->edgePtg);
            e.ptgPathIndex         = static_cast<int16_t>(n->edgeK);
            e.ptgTrimmableSpeed    = 1.0;  // never assigned by the reference's neighbour search: stays 1.0
            e.ptgFinalGoalRelSpeed = 0;
            e.ptgFinalRelativeGoal = goalRel;
            e.stateFrom            = par.state;  // frozen since the parent was popped
            e.stateTo              = n->state;
            e.ptgStep              = n->edgeStep;
            e.cost                 = n->edgeCost;
            edge_interpolated_path(e, trs, trs.ptgs[n->edgePtg]->getPathPose(n->edgeK, n->edgeStep), n->edgeStep, nSeg);
            tree.insert_child_of(*parentEdges, *parentNode, parentId, n->id.value(), n->firstState, std::move(e));
        }
        if (goalSnap) tree.node_state(goalSnap->first).pose = goalSnap->second;
    }
};

// ======================================================================================================
TPS_Astar::TPS_Astar(mppb_ctx* ctx) : impl_(new Impl())
{
    impl_->self = this;
    impl_->ctx  = ctx;
    heuristic   = [this](const SE2_KinState& from, const SE2orR2_KinState& goal) { return default_heuristic(from, goal); };
    impl_->defaultHeuristicType = &heuristic.target_type();  // (a user who assigns another callable changes this type)
}
TPS_Astar::~TPS_Astar()
{
    if (impl_->ownCtx && impl_->ctx) mppb_ctx_destroy(impl_->ctx);
}
mppb_ctx* TPS_Astar::context()
{
    impl_->ensureContext();
    return impl_->ctx;
}

void TPS_Astar::Impl::ensureContext()
{
    if (ctx) return;
    if (mppb_ctx_create(0, nullptr, &ctx) != MPPB_OK)
        throw std::runtime_error(std::string("TPS_Astar needs a CUDA device: ") + mppb_last_global_error());
    ownCtx = true;
}

void TPS_Astar::Impl::uploadPtgs(const TrajectoriesAndRobotShape& trs)
{
    std::vector<std::pair<const ptg_t*, uint64_t>> want;
    for (const auto& p : trs.ptgs) want.emplace_back(p.get(), p->revision());
    if (want == uploadedPtgs && ptgEpoch == mppb_ptgs_epoch(ctx) && static_cast<int>(want.size()) == mppb_num_ptgs(ctx)) return;
    ck(ctx, mppb_clear_ptgs(ctx));
    uploadedPtgs.clear();
    columnOffset.clear();
    for (size_t i = 0; i < trs.ptgs.size(); i++)
    {
        const ptg_t* p = trs.ptgs[i].get();
        if (auto* dd = dynamic_cast<const DiffDrive_C*>(p))
        {
            const mppb_ptg_grid_desc d = dd->gridDesc();
            ck(ctx, mppb_set_ptg_grid(ctx, static_cast<int>(i), &d));
            const mppb_ptg_paths_desc pd = dd->pathsDesc();
            ck(ctx, mppb_set_ptg_paths(ctx, static_cast<int>(i), &pd));
        }
        else if (auto* hb = dynamic_cast<const HolonomicBlend*>(p))
        {
            const mppb_ptg_holo_desc d = hb->holoDesc();
            ck(ctx, mppb_set_ptg_holo(ctx, static_cast<int>(i), &d));
        }
        else
            throw std::runtime_error("TPS_Astar: PTG #" + std::to_string(i) + " is of a kind the GPU library does not know");
        columnOffset.push_back(mppb_ptg_column_offset(ctx, static_cast<int>(i)));
    }
    uploadedPtgs = want;
    ptgEpoch     = mppb_ptgs_epoch(ctx);
}

namespace
{
// 64-bit content hash of a float array, four independent multiply-xor lanes (memory-bound: ~0.3 ms per 2^20 points)
uint64_t hash_floats(const float* p, size_t n, uint64_t seed)
{
    uint64_t     h[4] = {seed ^ 0x9E3779B97F4A7C15ull, seed ^ 0xC2B2AE3D27D4EB4Full, seed ^ 0x165667B19E3779F9ull, seed ^ 0x27D4EB2F165667C5ull};
    const size_t n8   = n / 8;
    for (size_t i = 0; i < n8; i++)
    {
        uint64_t w[4];
        std::memcpy(w, p + 8 * i, sizeof(w));
        for (int j = 0; j < 4; j++) h[j] = (h[j] ^ w[j]) * 0x100000001B3ull + (h[j] >> 29);
    }
    for (size_t i = 8 * n8; i < n; i++)
    {
        uint32_t w;
        std::memcpy(&w, p + i, sizeof(w));
        h[i & 3] = (h[i & 3] ^ w) * 0x100000001B3ull + (h[i & 3] >> 29);
    }
    uint64_t r = n;
    for (int j = 0; j < 4; j++) r = (r ^ h[j]) * 0xD6E8FEB86659FD93ull, r ^= r >> 32;
    return r;
}
}  // namespace

void TPS_Astar::Impl::uploadCloud(const std::vector<ObstacleSource::Ptr>& srcs, const float* box)
{
    CloudKey                     key;
    std::vector<PointCloud::Ptr> clouds;
    bool                         dynamic = false;
    for (const auto& os : srcs)
    {
        if (!os) continue;
        clouds.push_back(os->obstacles());
        dynamic = dynamic || os->dynamic();
        const PointCloud* c = clouds.back().get();
        if (c) MPP_ASSERT(c->xs.size() == c->ys.size());
        key.sizes.push_back(c ? c->size() : 0);
        key.hashes.push_back(c ? hash_floats(c->xs.data(), c->size(), 1) ^ hash_floats(c->ys.data(), c->size(), 2) : 0);
    }
    key.hasBox = box != nullptr;
    if (box) std::memcpy(key.box, box, sizeof(key.box));
    // still resident from the previous plan: same content, and nobody else touched the context's cloud since
    if (cloudValid && !dynamic && key == uploadedCloud && cloudEpoch == mppb_obstacles_epoch(ctx)) return;
    bool first = true;
    for (const auto& c : clouds)
    {
        static const float none = 0;
        const float*       xs   = c && c->size() ? c->xs.data() : &none;
        const float*       ys   = c && c->size() ? c->ys.data() : &none;
        ck(ctx, mppb_set_obstacles_clipped(ctx, xs, ys, c ? c->size() : 0, 0.0, first ? 0 : 1, box));
        first = false;
    }
    if (first) ck(ctx, mppb_set_obstacles_clipped(ctx, nullptr, nullptr, 0, 0.0, 0, box));
    uploadedCloud = key;
    cloudValid    = true;
    cloudEpoch    = mppb_obstacles_epoch(ctx);
}

void TPS_Astar::Impl::uploadEvaluators()
{
    // same evaluator objects with the same tables as in the previous plan: nothing to upload
    std::vector<EvalKey> want;
    for (const auto& cePtr : self->costEvaluators_)
    {
        const CostEvaluator* ce = cePtr.get();
        EvalKey              k;
        k.obj = ce;
        if (auto* cm = dynamic_cast<const CostEvaluatorCostMap*>(ce))
        {
            k.data = cm->cells.data();
            k.n    = cm->cells.size();
            k.p[0] = cm->x_min, k.p[1] = cm->y_min, k.p[2] = cm->params_.resolution, k.p[3] = cm->size_x;
            k.p[4] = cm->params_.useAverageOfPath ? 1 : 0;
        }
        else if (auto* wp = dynamic_cast<const CostEvaluatorPreferredWaypoint*>(ce))
        {
            k.data = wp->waypoints_.data();
            k.n    = wp->waypoints_.size();
            k.p[0] = wp->params_.waypointInfluenceRadius, k.p[1] = wp->params_.costScale;
            k.p[2] = wp->params_.useAverageOfPath ? 1 : 0;
            for (const auto& w : wp->waypoints_) k.p[3] += w.x, k.p[4] += w.y;  // cheap content check
        }
        want.push_back(k);
    }
    if (evalsValid && want == uploadedEvals && evalEpoch == mppb_evaluators_epoch(ctx)) return;
    evalsValid = false;
    ck(ctx, mppb_clear_evaluators(ctx));
    evalSlot.assign(self->costEvaluators_.size(), -1);
    numGpuEvals = 0;
    for (size_t e = 0; e < self->costEvaluators_.size(); e++)
    {
        const CostEvaluator* ce = self->costEvaluators_[e].get();
        MPP_ASSERT(ce);
        if (numGpuEvals >= MPPB_MAX_EVALUATORS) continue;  // the rest runs as host plugins
        if (auto* cm = dynamic_cast<const CostEvaluatorCostMap*>(ce))
        {
            mppb_costmap_desc d{cm->x_min, cm->y_min, cm->params_.resolution, cm->size_x, cm->size_y, cm->cells.data(),
                                cm->params_.useAverageOfPath ? 1 : 0};
            ck(ctx, mppb_set_costmap(ctx, numGpuEvals, &d));
            evalSlot[e] = numGpuEvals++;
        }
        else if (auto* wp = dynamic_cast<const CostEvaluatorPreferredWaypoint*>(ce))
        {
            std::vector<double> xs, ys;
            for (const auto& w : wp->waypoints_)
            {
                xs.push_back(w.x);
                ys.push_back(w.y);
            }
            ck(ctx, mppb_set_waypoints(ctx, numGpuEvals, xs.data(), ys.data(), xs.size(), wp->params_.waypointInfluenceRadius,
                                       wp->params_.costScale, wp->params_.useAverageOfPath ? 1 : 0));
            evalSlot[e] = numGpuEvals++;
        }
    }
    uploadedEvals = want;
    evalsValid    = true;
    evalEpoch     = mppb_evaluators_epoch(ctx);
}

// Decide whether this plan can use the device-side expansion and hand the planner parameters to the context.
void TPS_Astar::Impl::setupExpand()
{
    useExpand                = false;
    expandStride             = 0;
    static const bool disabled = std::getenv("MPPB_NO_EXPAND") != nullptr;  // A/B measurements and tests of the host phases
    if (disabled) return;
    for (int c : columnOffset)
        if (c < 0) return;  // a PTG of another kind (HolonomicBlend: its candidates need the host's expressions and libm)
    for (int slot : evalSlot)
        if (slot < 0) return;  // an evaluator that runs as a host plugin needs the edge on the host
    const TPS_Astar_Parameters& P = self->params_;
    if (P.ptg_sample_timestamps.empty() || P.ptg_sample_timestamps.size() > MPPB_MAX_TIMESTAMPS) return;
    mppb_astar_params a{};
    a.grid_resolution_xy              = P.grid_resolution_xy;
    a.grid_resolution_yaw             = P.grid_resolution_yaw;
    a.SE2_metricAngleWeight           = P.SE2_metricAngleWeight;
    a.heuristic_heading_weight        = P.heuristic_heading_weight;
    a.max_ptg_trajectories_to_explore = P.max_ptg_trajectories_to_explore;
    a.max_ptg_speeds_to_explore       = P.max_ptg_speeds_to_explore;
    a.path_interpolated_segments      = static_cast<uint32_t>(P.pathInterpolatedSegments);
    a.ptg_sample_timestamps           = P.ptg_sample_timestamps.data();
    a.num_timestamps                  = static_cast<uint32_t>(P.ptg_sample_timestamps.size());
    if (mppb_set_expand_params(ctx, &a) != MPPB_OK) return;
    expandStride = mppb_expand_max_neighbors(ctx);  // 0: outside the kernel's limits -> host phases
    useExpand    = expandStride > 0;
}

std::vector<PlannerOutput> TPS_Astar::Impl::run(const std::vector<const PlannerInput*>& ins, bool sharedCloud, int nThreads,
                                                bool materializeTrees)
{
    ensureContext();
    self->lastGpuStats = GpuStats();
    const size_t Q     = ins.size();
    std::vector<PlannerOutput> outs(Q);
    if (Q == 0) return outs;
    for (const PlannerInput* in : ins)
    {
        MPP_ASSERT(in->ptgs.initialized());
        if (in->ptgs.ptgs.size() != ins[0]->ptgs.ptgs.size()) throw std::runtime_error("plan_batch: all queries must use the same PTG set");
        for (size_t i = 0; i < in->ptgs.ptgs.size(); i++)
            if (in->ptgs.ptgs[i].get() != ins[0]->ptgs.ptgs[i].get())
                throw std::runtime_error("plan_batch: all queries must share the same PTG objects");
        if (sharedCloud && in->obstacles != ins[0]->obstacles)
            throw std::runtime_error("plan_batch: all queries must share the same obstacle sources");
    }
    auto tSetup = Clock::now();
    uploadPtgs(ins[0]->ptgs);
    uploadEvaluators();
    setupExpand();

    std::vector<std::shared_ptr<Query>> qs(Q);
    for (size_t i = 0; i < Q; i++)
    {
        qs[i]               = std::make_shared<Query>();
        Query& q            = *qs[i];
        q.planner           = self;
        q.P                 = &self->params_;
        q.in                = ins[i];
        q.columnOffset      = &columnOffset;
        q.trs               = ins[i]->ptgs;
        if (Q > 1)  // interleaved queries must not share mutable PTG state
            for (auto& p : q.trs.ptgs) p = std::shared_ptr<ptg_t>(p->cloneForQuery());
        q.deferTreeOps = Q == 1 && std::getenv("MPPB_NO_OVERLAP") == nullptr;
        q.useExpand    = useExpand;
        q.directHeuristic = defaultHeuristicType && self->heuristic && self->heuristic.target_type() == *defaultHeuristicType;
        q.lazyTree     = useExpand && !materializeTrees;
        q.start();
    }
    // obstacles: one query -> clipped to its world box at upload (ObstacleSource::apply_clipping_box);
    // many queries -> the shared cloud stays whole and each node carries its query's box
    if (sharedCloud)
        uploadCloud(ins[0]->obstacles, nullptr);
    else
        uploadCloud(ins[0]->obstacles, qs[0]->box);

    const int       hw = static_cast<int>(std::max(1u, std::thread::hardware_concurrency()));
    Workers         pool(Q == 1 ? 1 : (nThreads > 0 ? nThreads : hw));
    const int       cols = mppb_total_paths(ctx);
    std::vector<size_t> active, holoBase, edgeBase;
    active.reserve(Q);
    GpuStats& st = self->lastGpuStats;

    auto lap = [&st](Clock::time_point& t, int phase) {
        const auto now = Clock::now();
        st.phaseSeconds[phase] += std::chrono::duration<double>(now - t).count();
        t = now;
    };
    lap(tSetup, 0);
    // ---- device-side expansion: per round ONE parallel host section (relaxation of the previous round's neighbours,
    // pop, goal direction per PTG) and ONE GPU call (collision rows + expansion of every current node). A batch is
    // cut into two waves that alternate: while the GPU expands the nodes of one wave, the workers relax the other
    // wave's neighbours, so the GPU time (a quarter of a round on a 16-core host) costs no wall time. ----
    if (useExpand)
    {
        struct Wave
        {
            std::vector<size_t>      members, active;
            Pinned<mppb_expand_node> in;
            Pinned<mppb_neighbor>    nbrs;
            Pinned<uint32_t>         counts;
            bool                     inFlight = false;
        };
        static const bool   noWaves = std::getenv("MPPB_NO_WAVES") != nullptr;  // A/B measurements only
        static const size_t minQ    = std::getenv("MPPB_WAVES_MIN_QUERIES") ? std::strtoul(std::getenv("MPPB_WAVES_MIN_QUERIES"), nullptr, 10)
                                                                             : 64;  // (tests lower it)
        const size_t        nW      = (Q >= std::max<size_t>(2, minQ) && !noWaves) ? 2 : 1;
        Wave              W[2];
        for (size_t i = 0; i < Q; i++) W[i % nW].members.push_back(i);
        const double clip = qs[0]->MAX_XY_DIST;
        // relaxation of the wave's previous neighbours, pop, phase A; then the wave's input records
        auto hostSection = [&](Wave& w) {
            auto                 tl     = Clock::now();
            const mppb_neighbor* nbrs   = w.nbrs.p;
            const uint32_t*      counts = w.counts.p;
            pool.run(w.members.size(), [&](size_t j) {
                if (kPrefetchNextQuery)
                {
                    // a worker walks a contiguous slice of queries whose search states are cold again by the time their
                    // turn comes: the query after next gets its hot fields requested, the next one (whose fields were
                    // requested a turn ago) its neighbour records and current node
                    const size_t nM = w.members.size();
                    if (j + 2 < nM)
                    {
                        const Query* q2 = qs[w.members[j + 2]].get();
                        __builtin_prefetch(&q2->grid);
                        __builtin_prefetch(&q2->current);
                        __builtin_prefetch(&q2->po.nodesExpanded);
                    }
                    if (j + 1 < nM)
                    {
                        const Query* q1 = qs[w.members[j + 1]].get();
                        if (q1->needF)
                        {
                            const char* r = reinterpret_cast<const char*>(nbrs + q1->slot * expandStride);
                            for (int l = 0; l < 4; l++) __builtin_prefetch(r + 64 * l);
                            __builtin_prefetch(counts + 2 * q1->slot);
                        }
                        if (q1->current) __builtin_prefetch(q1->current);
                    }
                }
                Query& q = *qs[w.members[j]];
                q.active = false;
                if (q.done) return;
                if (q.needF)
                {
                    q.consumeNeighbors(nbrs + q.slot * expandStride, counts[2 * q.slot], counts[2 * q.slot + 1]);
                    q.needF = false;
                }
                {
                    MPPB_PROF(0);
                    q.active = q.popNext();
                }
                if (q.active)
                {
                    MPPB_PROF(1);
                    q.phaseAExpand();
                }
            });
            w.active.clear();
            for (size_t i : w.members)
                if (qs[i]->active) w.active.push_back(i);
            const size_t A = w.active.size();
            if (A)
            {
                mppb_expand_node* in = w.in.reserve(A);
                w.nbrs.reserve(A * expandStride);
                w.counts.reserve(2 * A);
                for (size_t a = 0; a < A; a++)
                {
                    Query& q = *qs[w.active[a]];
                    q.slot   = a;
                    in[a]    = q.expIn;
                }
            }
            lap(tl, 1);
        };
        auto launch = [&](Wave& w) {
            const size_t A = w.active.size();
            if (!A) return;
            ck(ctx, mppb_expand_nodes_begin(ctx, A, w.in.p, sharedCloud ? 1 : 0, clip, w.nbrs.p, w.counts.p));
            w.inFlight = true;
            st.collisionBatches++;
            st.nodesEvaluated += A;
        };
        auto wait = [&](Wave& w) {
            if (!w.inFlight) return;
            auto tl = Clock::now();
            ck(ctx, mppb_sync(ctx));
            w.inFlight = false;
            for (size_t a = 0; a < w.active.size(); a++)
            {
                qs[w.active[a]]->needF = true;
                st.edgesCosted += w.counts.p[2 * a];
            }
            st.gpuSeconds += seconds_since(tl);  // the part of the GPU call the host could not hide
            lap(tl, 2);
        };
        hostSection(W[0]);
        launch(W[0]);
        if (nW == 2) hostSection(W[1]);
        for (size_t turn = 0;;)
        {
            // here: W[turn] is on the GPU (or has nothing left); with two waves, the other one is ready to launch
            wait(W[turn]);
            if (nW == 2) launch(W[1 - turn]);
            hostSection(W[turn]);
            if (nW == 1)
            {
                if (W[0].active.empty()) break;
                launch(W[0]);
                continue;
            }
            if (W[0].active.empty() && W[1].active.empty() && !W[0].inFlight && !W[1].inFlight) break;
            turn = 1 - turn;
        }
    }
    double* costs = nullptr;  // evaluator values of the previous round's tentative edges (pinned buffer)
    for (; !useExpand;)
    {
        auto tl = Clock::now();
        // ---- per query, no barrier in between: F of the previous round, pop, A of this round ----
        // A single query overlaps the second half of A (end poses and lattice cells of the candidate TPS points, host)
        // with phase B (GPU): the collision calls are issued in their split form right after phaseAPrepare() and
        // completed once phaseACandidates() is done. A batch of queries keeps A whole inside the parallel section.
        static const bool noOverlap = std::getenv("MPPB_NO_OVERLAP") != nullptr;  // A/B measurements only
        const bool        overlapAB = Q == 1 && !noOverlap;
        pool.run(Q, [&](size_t i) {
            Query& q = *qs[i];
            q.active = false;
            if (q.done) return;
            if (q.needF)
            {
                q.phaseF(costs ? costs + q.costBase * numGpuEvals : nullptr, numGpuEvals, evalSlot);
                q.needF = false;
            }
            q.active = q.popNext();
            if (q.active)
            {
                q.phaseAPrepare();
                if (!overlapAB) q.phaseACandidates();
            }
        });
        active.clear();
        for (size_t i = 0; i < Q; i++)
            if (qs[i]->active)
            {
                qs[i]->slot = active.size();
                active.push_back(i);
            }
        if (active.empty()) break;
        const size_t A = active.size();
        lap(tl, 1);

        // ---- phase B: one collision batch for all current nodes ----
        double* poses = hPoses.reserve(3 * A);
        float*  boxes = sharedCloud ? hBoxes.reserve(4 * A) : nullptr;
        holoBase.assign(A + 1, 0);
        for (size_t a = 0; a < A; a++)
        {
            const Query& q   = *qs[active[a]];
            poses[3 * a + 0] = q.current->state.pose.x;
            poses[3 * a + 1] = q.current->state.pose.y;
            poses[3 * a + 2] = q.current->state.pose.phi;
            if (boxes) std::memcpy(boxes + 4 * a, q.box, sizeof(q.box));
            holoBase[a + 1] = holoBase[a] + q.holoCands.size();
        }
        const auto tB   = Clock::now();
        float*     rows = nullptr;
        if (cols > 0)
        {
            rows = hRows.reserve(static_cast<size_t>(cols) * A);
            if (overlapAB)
                ck(ctx, mppb_eval_nodes_boxed_begin(ctx, A, poses, boxes, qs[active[0]]->MAX_XY_DIST, rows));
            else
                ck(ctx, mppb_eval_nodes_boxed(ctx, A, poses, boxes, qs[active[0]]->MAX_XY_DIST, rows));
        }
        double* holoOut = nullptr;
        if (holoBase[A] > 0)
        {
            mppb_holo_cand* hc = hCands.reserve(holoBase[A]);
            for (size_t a = 0; a < A; a++)
            {
                const auto& v = qs[active[a]]->holoCands;
                if (v.empty()) continue;
                std::memcpy(hc + holoBase[a], v.data(), v.size() * sizeof(mppb_holo_cand));
                for (size_t j = 0; j < v.size(); j++) hc[holoBase[a] + j].node = static_cast<int32_t>(a);
            }
            holoOut = hHoloOut.reserve(holoBase[A]);
            if (overlapAB)
                ck(ctx, mppb_eval_holo_candidates_boxed_begin(ctx, A, poses, boxes, holoBase[A], hc, qs[active[0]]->MAX_XY_DIST, holoOut));
            else
                ck(ctx, mppb_eval_holo_candidates_boxed(ctx, A, poses, boxes, holoBase[A], hc, qs[active[0]]->MAX_XY_DIST, holoOut));
        }
        if (overlapAB)
        {
            qs[active[0]]->phaseACandidates();  // host work while the GPU evaluates
            qs[active[0]]->flushTreeOps();      // (the previous expansion's tree insertions)
            ck(ctx, mppb_sync(ctx));
        }
        st.gpuSeconds += seconds_since(tB);
        st.collisionBatches++;
        st.nodesEvaluated += A;
        st.holoCandidates += holoBase[A];
        lap(tl, 2);

        // ---- phases C + D ----
        pool.run(Q, [&](size_t i) {  // over the queries, not the active list: a query stays with its worker
            Query& q = *qs[i];
            if (!q.active) return;
            const size_t a = q.slot;
            q.phaseCD(rows ? rows + static_cast<size_t>(cols) * a : nullptr, static_cast<uint32_t>(cols),
                      holoOut ? holoOut + holoBase[a] : nullptr);
            q.needF = true;
        });
        lap(tl, 3);

        // ---- phase E: one cost batch for all tentative edges ----
        edgeBase.assign(A + 1, 0);
        for (size_t a = 0; a < A; a++)
        {
            qs[active[a]]->costBase = edgeBase[a];
            edgeBase[a + 1]         = edgeBase[a] + qs[active[a]]->pending.size();
        }
        const size_t E = edgeBase[A];
        costs          = nullptr;
        if (E > 0 && numGpuEvals > 0)
        {
            size_t nSamples = 0;
            for (size_t a = 0; a < A; a++)
                for (const auto& pe : qs[active[a]]->pending) nSamples += pe.edge.interpolatedPath.size();
            double*   from = hFrom.reserve(3 * E);
            uint32_t* off  = hOffsets.reserve(E + 1);
            double*   rel  = hRel.reserve(2 * nSamples);
            costs          = hCost.reserve(E * numGpuEvals);
            size_t e = 0, s = 0;
            off[0] = 0;
            for (size_t a = 0; a < A; a++)
                for (const auto& pe : qs[active[a]]->pending)
                {
                    from[3 * e + 0] = pe.edge.stateFrom.pose.x;
                    from[3 * e + 1] = pe.edge.stateFrom.pose.y;
                    from[3 * e + 2] = pe.edge.stateFrom.pose.phi;
                    for (const auto& kv : pe.edge.interpolatedPath)
                    {
                        rel[2 * s + 0] = kv.second.x;
                        rel[2 * s + 1] = kv.second.y;
                        s++;
                    }
                    off[++e] = static_cast<uint32_t>(s);
                }
            lap(tl, 4);
            const auto tE = Clock::now();
            ck(ctx, mppb_eval_edge_costs(ctx, E, from, off, rel, costs));
            st.gpuSeconds += seconds_since(tE);
            st.costBatches++;
            st.edgesCosted += E;
            lap(tl, 5);
        }
        // phase F (relaxation) of these edges runs at the head of the next round, fused with pop and A
    }
    // hand the results over and tear the per-query search state down in parallel (millions of small nodes)
    auto tEnd = Clock::now();
    for (size_t i = 0; i < Q; i++) st.holoNearTies += qs[i]->nearTies;
    pool.run(Q, [&](size_t i) {
        outs[i] = std::move(qs[i]->po);
        if (qs[i]->lazyTree)
        {
            // the query's search state stays alive behind the output until the tree has been built from it (or the
            // output dies); what the builder reads no longer refers to the planner or to the caller's input
            std::shared_ptr<Query> q = std::move(qs[i]);
            q->planner               = nullptr;
            q->P                     = nullptr;
            q->in                    = nullptr;
            outs[i].treeBuilder      = [q](PlannerOutput& o) { q->buildTree(o.motionTree, o.originalInput); };
        }
        qs[i].reset();
    });
    lap(tEnd, 6);
    // The reference's profiler sections (TPS_Astar.cpp:89 plan, :195 plan.iter, :548 find_feasible, :574
    // find_feasible.ptgUpdateDyn, :605 find_feasible.loop1, :701 find_feasible.loop2, :740
    // find_feasible.tp_obstacles_single, :803 find_feasible.finalFill, :833 cached_local_obstacles), fed from the
    // phase timers of this restructured loop. One expansion = one plan.iter = one find_feasible. The clip/transform
    // of cached_local_obstacles is fused into the collision kernels: it is counted, its time is part of
    // find_feasible.tp_obstacles_single. loop1 + ptgUpdateDyn (candidate enumeration) share phase A's timer with the
    // relaxation of the previous round; loop2 is the blocked test + selection (phases C, D; on the device-expansion
    // path inside the GPU call); finalFill is part of the same.
    {
        const double* ph = st.phaseSeconds;
        size_t        expansions = 0, tps = 0;
        for (const auto& o : outs)
        {
            expansions += o.nodesExpanded;
            tps += o.tpsPointsEvaluated;
        }
        TimeLogger&  prof = self->profiler_();
        const double loop = ph[1] + ph[2] + ph[3] + ph[4] + ph[5];
        prof.registerUserMeasure("plan", ph[0] + loop + ph[6], Q);
        prof.registerUserMeasure("plan.iter", loop, expansions);
        prof.registerUserMeasure("find_feasible", ph[1] + ph[2] + ph[3], expansions);
        prof.registerUserMeasure("find_feasible.loop1", ph[1], expansions);
        prof.registerUserMeasure("find_feasible.ptgUpdateDyn", 0.0, expansions * ins[0]->ptgs.ptgs.size());
        prof.registerUserMeasure("find_feasible.loop2", ph[3], expansions);
        prof.registerUserMeasure("find_feasible.tp_obstacles_single", ph[2], tps);
        prof.registerUserMeasure("find_feasible.finalFill", 0.0, expansions);
        prof.registerUserMeasure("cached_local_obstacles", 0.0, st.nodesEvaluated);
    }
    return outs;
}

PlannerOutput TPS_Astar::plan(const PlannerInput& in)
{
    std::vector<const PlannerInput*> v{&in};
    return std::move(impl_->run(v, false, 1, true)[0]);
}

std::vector<PlannerOutput> TPS_Astar::plan_batch(const std::vector<PlannerInput>& ins, int nThreads, bool materializeTrees)
{
    std::vector<const PlannerInput*> v;
    for (const auto& i : ins) v.push_back(&i);
    return impl_->run(v, true, nThreads, materializeTrees);
}

// ---- heuristics (TPS_Astar.cpp:477-537) --------------------------------------------------------------
cost_t TPS_Astar::default_heuristic_SE2(const SE2_KinState& from, const TPose2D& goal) const
{
    const TPose2D relPose     = goal - from.pose;
    const double  distSE2     = relPose.norm() + params_.SE2_metricAngleWeight * std::abs(relPose.phi);
    const double  distHeading = (relPose.norm() < 0.1) ? 0.0 : std::abs(ang_distance(std::atan2(relPose.y, relPose.x), from.pose.phi));
    return distSE2 + params_.heuristic_heading_weight * distHeading;
}
cost_t TPS_Astar::default_heuristic_R2(const SE2_KinState& from, const TPoint2D& goal) const
{
    // `goal - from.pose` with a point goal: plain vector difference (MRPT ambiguity)
    const double dx = from.pose.x - goal.x, dy = from.pose.y - goal.y;
    const double distR2 = std::sqrt(dx * dx + dy * dy);
    const double rx = goal.x - from.pose.x, ry = goal.y - from.pose.y;
    const double distHeading = (std::sqrt(rx * rx + ry * ry) < 0.1) ? 0.0 : std::abs(ang_distance(std::atan2(ry, rx), from.pose.phi));
    return distR2 + params_.heuristic_heading_weight * distHeading;
}
cost_t TPS_Astar::default_heuristic(const SE2_KinState& from, const SE2orR2_KinState& goal) const
{
    if (goal.state.isPoint()) return default_heuristic_R2(from, goal.state.point());
    if (goal.state.isPose()) return default_heuristic_SE2(from, goal.state.pose());
    throw std::runtime_error("Goal of unknown type?");
}

// ---- parameters (TPS_Astar.cpp:26-79) ------------------------------------------------------------------
void TPS_Astar_Parameters::load_from_yaml(const Yaml& c)
{
    MPP_ASSERT(c.isMap());
    grid_resolution_xy              = c.getDouble("grid_resolution_xy");
    grid_resolution_yaw             = c.getDouble("grid_resolution_yaw") * M_PI / 180.0;
    SE2_metricAngleWeight           = c.getDouble("SE2_metricAngleWeight");
    max_ptg_trajectories_to_explore = static_cast<uint32_t>(c.getDouble("max_ptg_trajectories_to_explore"));
    max_ptg_speeds_to_explore       = static_cast<uint32_t>(c.getDouble("max_ptg_speeds_to_explore"));
    ptg_sample_timestamps           = c.getDoubleVector("ptg_sample_timestamps");
    pathInterpolatedSegments        = c.getOrDefault<size_t>("pathInterpolatedSegments", pathInterpolatedSegments);
    heuristic_heading_weight        = c.getOrDefault<double>("heuristic_heading_weight", heuristic_heading_weight);
    maximumComputationTime          = c.getOrDefault<double>("maximumComputationTime", maximumComputationTime);
    if (c.has("maxExpansions")) maxExpansions = static_cast<size_t>(c.getDouble("maxExpansions"));
}
TPS_Astar_Parameters TPS_Astar_Parameters::FromYAML(const Yaml& c)
{
    TPS_Astar_Parameters p;
    p.load_from_yaml(c);
    return p;
}
std::string TPS_Astar_Parameters::as_yaml_text() const
{
    std::stringstream s;
    s << "SE2_metricAngleWeight: " << SE2_metricAngleWeight << "\n"
      << "pathInterpolatedSegments: " << pathInterpolatedSegments << "\n"
      << "grid_resolution_xy: " << grid_resolution_xy << "\n"
      << "heuristic_heading_weight: " << heuristic_heading_weight << "\n"
      << "max_ptg_trajectories_to_explore: " << max_ptg_trajectories_to_explore << "\n"
      << "max_ptg_speeds_to_explore: " << max_ptg_speeds_to_explore << "\n"
      << "grid_resolution_yaw: " << grid_resolution_yaw * 180.0 / M_PI << "\n"
      << "maximumComputationTime: " << maximumComputationTime << "\n"
      << "ptg_sample_timestamps: [";
    for (size_t i = 0; i < ptg_sample_timestamps.size(); i++) s << (i ? ", " : "") << ptg_sample_timestamps[i];
    s << "]\n";
    return s.str();
}

}  // namespace mpp
