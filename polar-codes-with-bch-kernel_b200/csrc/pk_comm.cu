// pk_comm.cu -- multi-GPU part of the C ABI: frames of one SNR point sharded over GPUs, counters combined by
// ONE NCCL all-reduce per point (the reference has a single sequential loop, src/dataForPlot.cpp:41-95).
//
// Two ways to span the GPUs of a box, same entry points:
//   * one process, ndev devices (pk_comm_create): ncclCommInitAll, one stream per device, group calls;
//   * one process per GPU (pk_comm_create_rank; what `torchrun` / mpirun start): ncclCommInitRank with a unique id the
//     caller passes between its processes (pk_comm_unique_id on rank 0).
// NCCL is bound at run time (dlopen of libnccl.so.2) so that libpkb200.so carries no link-time dependency on it:
// inside a PyTorch process the already loaded NCCL is found, elsewhere the system one.  With one rank no NCCL is needed.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/pk_capi.h"

int pk_set_error(int code, const std::string &msg);   // pk_capi.cu
extern unsigned long long g_pk_launches;

namespace {

struct NcclApi {
    void *lib = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommInitAll) CommInitAll = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    std::string err;
};

NcclApi &nccl() {
    static NcclApi api;
    if (api.lib || !api.err.empty()) return api;
    for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
        api.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
        if (api.lib) break;
    }
    if (!api.lib) {
        api.err = std::string("NCCL not found (dlopen libnccl.so.2): ") + (dlerror() ? dlerror() : "");
        return api;
    }
#define PK_SYM(field, sym)                                                       \
    api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.lib, #sym));      \
    if (!api.field) api.err = "NCCL symbol missing: " #sym;
    PK_SYM(GetUniqueId, ncclGetUniqueId)
    PK_SYM(CommInitRank, ncclCommInitRank)
    PK_SYM(CommInitAll, ncclCommInitAll)
    PK_SYM(CommDestroy, ncclCommDestroy)
    PK_SYM(AllReduce, ncclAllReduce)
    PK_SYM(GroupStart, ncclGroupStart)
    PK_SYM(GroupEnd, ncclGroupEnd)
    PK_SYM(GetErrorString, ncclGetErrorString)
#undef PK_SYM
    return api;
}

// pk_point_result (8 x u64: six sums, a maximum, an OR of flag bits) <-> the vector that is all-reduced:
// [0..6) sums, [6..14) one count per flag bit, [14] maximum (reduced on its own with ncclMax)
enum { PK_PACK_SUMS = 14, PK_PACK_LEN = 16 };
__global__ void k_point_pack(const unsigned long long *res, unsigned long long *pack) {
    const int i = threadIdx.x;
    if (i < 6) pack[i] = res[i];
    else if (i < 14) pack[i] = (res[7] >> (i - 6)) & 1ull;
    else if (i == 14) pack[14] = res[6];
    else pack[15] = 0;
}
__global__ void k_point_unpack(const unsigned long long *pack, unsigned long long *res) {
    const int i = threadIdx.x;
    if (i < 6) res[i] = pack[i];
    else if (i == 6) res[6] = pack[14];
    else if (i == 7) {
        unsigned long long f = 0;
        for (int b = 0; b < 8; ++b) f |= (pack[6 + b] ? 1ull : 0ull) << b;
        res[7] = f;
    }
}

}  // namespace

struct pk_comm {
    int world = 1, rank0 = 0;            // ranks of the whole job, global rank of local device 0
    std::vector<int> dev;                // local device ordinals (single process: all of them; rank mode: one)
    std::vector<ncclComm_t> comms;       // one per local device (empty when world == 1)
    std::vector<cudaStream_t> streams;
    std::vector<unsigned long long *> d_pack;   // PK_PACK_LEN per local device
    std::vector<unsigned long long *> d_vec;    // `world` slots per local device (pk_comm_allreduce_u64)
};

struct pk_comm_kaneko {
    pk_comm *comm = nullptr;
    std::vector<pk_code *> codes;
    std::vector<pk_kaneko *> decs;
    std::vector<unsigned long long *> d_tot;    // pk_point_result per local device
    std::vector<pk_frame_rec *> d_rec, h_rec;   // per-frame records of a round (finite e)
    long rec_cap = 0;
    int n = 0;
};

#define PKC_CUDA(expr)                                                                              \
    do {                                                                                            \
        cudaError_t e__ = (expr);                                                                   \
        if (e__ != cudaSuccess) return pk_set_error(PK_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__)); \
    } while (0)
#define PKC_NCCL(expr)                                                                              \
    do {                                                                                            \
        ncclResult_t r__ = (expr);                                                                  \
        if (r__ != ncclSuccess) return pk_set_error(PK_ERR_CUDA, std::string(#expr) + ": " + nccl().GetErrorString(r__)); \
    } while (0)

static int comm_alloc(pk_comm *c) {
    for (size_t i = 0; i < c->dev.size(); ++i) {
        PKC_CUDA(cudaSetDevice(c->dev[i]));
        cudaStream_t st = nullptr;
        PKC_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        c->streams.push_back(st);
        unsigned long long *p = nullptr, *v = nullptr;
        PKC_CUDA(cudaMalloc(&p, PK_PACK_LEN * sizeof(unsigned long long)));
        c->d_pack.push_back(p);
        PKC_CUDA(cudaMalloc(&v, (size_t)std::max(c->world, 1) * sizeof(unsigned long long)));
        c->d_vec.push_back(v);
    }
    return PK_OK;
}

extern "C" {

int pk_comm_create(int ndev, const int *devices, pk_comm **out) {
    if (!out) return pk_set_error(PK_ERR_ARG, "out is NULL");
    *out = nullptr;
    int have = 0;
    if (cudaGetDeviceCount(&have) != cudaSuccess || have <= 0) return pk_set_error(PK_ERR_CUDA, "no CUDA device: libpkb200 has no CPU path");
    if (ndev < 1 || ndev > have) return pk_set_error(PK_ERR_ARG, "pk_comm_create: ndev must be in [1, device count]");
    pk_comm *c = new (std::nothrow) pk_comm;
    if (!c) return pk_set_error(PK_ERR_ALLOC, "out of memory");
    c->world = ndev;
    for (int i = 0; i < ndev; ++i) {
        const int d = devices ? devices[i] : i;
        if (d < 0 || d >= have) { delete c; return pk_set_error(PK_ERR_ARG, "bad device ordinal"); }
        c->dev.push_back(d);
    }
    int rc = comm_alloc(c);
    if (rc == PK_OK && ndev > 1) {
        NcclApi &api = nccl();
        if (!api.err.empty()) rc = pk_set_error(PK_ERR_UNSUPPORTED, api.err);
        else {
            c->comms.resize(ndev);
            ncclResult_t r = api.CommInitAll(c->comms.data(), ndev, c->dev.data());
            if (r != ncclSuccess) { c->comms.clear(); rc = pk_set_error(PK_ERR_CUDA, std::string("ncclCommInitAll: ") + api.GetErrorString(r)); }
        }
    }
    if (rc != PK_OK) { pk_comm_destroy(c); return rc; }
    *out = c;
    return PK_OK;
}

int pk_comm_unique_id(uint8_t *id128) {
    if (!id128) return pk_set_error(PK_ERR_ARG, "NULL");
    NcclApi &api = nccl();
    if (!api.err.empty()) return pk_set_error(PK_ERR_UNSUPPORTED, api.err);
    ncclUniqueId id;
    PKC_NCCL(api.GetUniqueId(&id));
    static_assert(sizeof(id) == 128, "ncclUniqueId is 128 bytes");
    std::memcpy(id128, &id, 128);
    return PK_OK;
}

int pk_comm_create_rank(int world, int rank, const uint8_t *id128, int device, pk_comm **out) {
    if (!out) return pk_set_error(PK_ERR_ARG, "out is NULL");
    *out = nullptr;
    if (world < 1 || rank < 0 || rank >= world || (world > 1 && !id128)) return pk_set_error(PK_ERR_ARG, "pk_comm_create_rank: bad world / rank / id");
    int have = 0;
    if (cudaGetDeviceCount(&have) != cudaSuccess || have <= 0) return pk_set_error(PK_ERR_CUDA, "no CUDA device: libpkb200 has no CPU path");
    if (device < 0 || device >= have) return pk_set_error(PK_ERR_ARG, "bad device ordinal");
    pk_comm *c = new (std::nothrow) pk_comm;
    if (!c) return pk_set_error(PK_ERR_ALLOC, "out of memory");
    c->world = world;
    c->rank0 = rank;
    c->dev.push_back(device);
    int rc = comm_alloc(c);
    if (rc == PK_OK && world > 1) {
        NcclApi &api = nccl();
        if (!api.err.empty()) rc = pk_set_error(PK_ERR_UNSUPPORTED, api.err);
        else {
            ncclUniqueId id;
            std::memcpy(&id, id128, 128);
            ncclComm_t cm = nullptr;
            cudaSetDevice(device);
            ncclResult_t r = api.CommInitRank(&cm, world, id, rank);
            if (r != ncclSuccess) rc = pk_set_error(PK_ERR_CUDA, std::string("ncclCommInitRank: ") + api.GetErrorString(r));
            else c->comms.push_back(cm);
        }
    }
    if (rc != PK_OK) { pk_comm_destroy(c); return rc; }
    *out = c;
    return PK_OK;
}

void pk_comm_destroy(pk_comm *c) {
    if (!c) return;
    for (size_t i = 0; i < c->dev.size(); ++i) {
        cudaSetDevice(c->dev[i]);
        if (i < c->streams.size() && c->streams[i]) { cudaStreamSynchronize(c->streams[i]); }
        if (i < c->comms.size() && c->comms[i]) nccl().CommDestroy(c->comms[i]);
        if (i < c->streams.size() && c->streams[i]) cudaStreamDestroy(c->streams[i]);
        if (i < c->d_pack.size()) cudaFree(c->d_pack[i]);
        if (i < c->d_vec.size()) cudaFree(c->d_vec[i]);
    }
    delete c;
}

int pk_comm_size(const pk_comm *c) { return c ? c->world : 0; }
int pk_comm_rank(const pk_comm *c) { return c ? c->rank0 : -1; }
int pk_comm_local_devices(const pk_comm *c) { return c ? (int)c->dev.size() : 0; }
void *pk_comm_stream(const pk_comm *c, int local) { return (c && local >= 0 && local < (int)c->streams.size()) ? (void *)c->streams[local] : nullptr; }

// Enqueues, on the communicator's stream of every local device, the in-place reduction over all ranks of the
// pk_point_result at d_results[i] (device memory of local device i): sums of the six counters, maximum of
// max_trials_seen, OR of flags_or.  Asynchronous; pk_comm_sync waits.
int pk_allreduce_point(pk_comm *c, pk_point_result *const *d_results) {
    if (!c || !d_results) return pk_set_error(PK_ERR_ARG, "NULL argument");
    if (c->world == 1) return PK_OK;
    NcclApi &api = nccl();
    const size_t nloc = c->dev.size();
    for (size_t i = 0; i < nloc; ++i) {
        PKC_CUDA(cudaSetDevice(c->dev[i]));
        k_point_pack<<<1, 16, 0, c->streams[i]>>>(reinterpret_cast<const unsigned long long *>(d_results[i]), c->d_pack[i]);
        ++g_pk_launches;
        PKC_CUDA(cudaGetLastError());
    }
    PKC_NCCL(api.GroupStart());
    for (size_t i = 0; i < nloc; ++i) {
        PKC_NCCL(api.AllReduce(c->d_pack[i], c->d_pack[i], PK_PACK_SUMS, ncclUint64, ncclSum, c->comms[i], c->streams[i]));
        PKC_NCCL(api.AllReduce(c->d_pack[i] + 14, c->d_pack[i] + 14, 1, ncclUint64, ncclMax, c->comms[i], c->streams[i]));
    }
    PKC_NCCL(api.GroupEnd());
    for (size_t i = 0; i < nloc; ++i) {
        PKC_CUDA(cudaSetDevice(c->dev[i]));
        k_point_unpack<<<1, 16, 0, c->streams[i]>>>(c->d_pack[i], reinterpret_cast<unsigned long long *>(d_results[i]));
        ++g_pk_launches;
        PKC_CUDA(cudaGetLastError());
    }
    return PK_OK;
}

int pk_comm_sync(pk_comm *c) {
    if (!c) return pk_set_error(PK_ERR_ARG, "NULL");
    for (size_t i = 0; i < c->dev.size(); ++i) {
        PKC_CUDA(cudaSetDevice(c->dev[i]));
        PKC_CUDA(cudaStreamSynchronize(c->streams[i]));
    }
    return PK_OK;
}

// Host vector of `world` u64 slots, summed over ranks in place (every rank fills its own slot(s), zero elsewhere):
// the all-gather of the per-rank frame-error counts that the stop rule needs (rank mode).
static int allreduce_host_vec(pk_comm *c, std::vector<unsigned long long> &v) {
    if (c->world == 1 || c->dev.size() != 1) return PK_OK;
    NcclApi &api = nccl();
    PKC_CUDA(cudaSetDevice(c->dev[0]));
    PKC_CUDA(cudaMemcpyAsync(c->d_vec[0], v.data(), v.size() * 8, cudaMemcpyHostToDevice, c->streams[0]));
    PKC_NCCL(api.AllReduce(c->d_vec[0], c->d_vec[0], v.size(), ncclUint64, ncclSum, c->comms[0], c->streams[0]));
    PKC_CUDA(cudaMemcpyAsync(v.data(), c->d_vec[0], v.size() * 8, cudaMemcpyDeviceToHost, c->streams[0]));
    PKC_CUDA(cudaStreamSynchronize(c->streams[0]));
    return PK_OK;
}

// ------------------------------------------------------------------ sharded Kaneko decoder
int pk_comm_kaneko_create(pk_comm *c, int m, int t, double llr_snr_db, long J, long max_trials, pk_comm_kaneko **out) {
    if (!c || !out) return pk_set_error(PK_ERR_ARG, "NULL argument");
    *out = nullptr;
    pk_comm_kaneko *k = new (std::nothrow) pk_comm_kaneko;
    if (!k) return pk_set_error(PK_ERR_ALLOC, "out of memory");
    k->comm = c;
    int rc = PK_OK;
    for (size_t i = 0; i < c->dev.size() && rc == PK_OK; ++i) {
        pk_code *code = nullptr;
        pk_kaneko *dec = nullptr;
        rc = pk_code_create(m, t, c->dev[i], &code);
        if (rc == PK_OK) { k->codes.push_back(code); rc = pk_kaneko_create(code, llr_snr_db, J, max_trials, &dec); }
        if (rc == PK_OK) {
            k->decs.push_back(dec);
            unsigned long long *p = nullptr;
            if (cudaSetDevice(c->dev[i]) != cudaSuccess || cudaMalloc(&p, sizeof(pk_point_result)) != cudaSuccess)
                rc = pk_set_error(PK_ERR_CUDA, "pk_comm_kaneko_create: cudaMalloc failed");
            else k->d_tot.push_back(p);
        }
    }
    if (rc == PK_OK) pk_code_info(k->codes[0], &k->n, nullptr, nullptr, nullptr, nullptr);
    if (rc != PK_OK) { pk_comm_kaneko_destroy(k); return rc; }
    *out = k;
    return PK_OK;
}

void pk_comm_kaneko_destroy(pk_comm_kaneko *k) {
    if (!k) return;
    for (size_t i = 0; i < k->comm->dev.size(); ++i) {
        cudaSetDevice(k->comm->dev[i]);
        if (i < k->d_tot.size()) cudaFree(k->d_tot[i]);
        if (i < k->d_rec.size()) cudaFree(k->d_rec[i]);
        if (i < k->h_rec.size() && k->h_rec[i]) cudaFreeHost(k->h_rec[i]);
    }
    for (pk_kaneko *d : k->decs) pk_kaneko_destroy(d);
    for (pk_code *cd : k->codes) pk_code_destroy(cd);
    delete k;
}

pk_kaneko *pk_comm_kaneko_local(pk_comm_kaneko *k, int local) { return (k && local >= 0 && local < (int)k->decs.size()) ? k->decs[local] : nullptr; }

static int ensure_recs(pk_comm_kaneko *k, long cap) {
    if (cap <= k->rec_cap) return PK_OK;
    const size_t nloc = k->comm->dev.size();
    for (size_t i = 0; i < nloc; ++i) {
        PKC_CUDA(cudaSetDevice(k->comm->dev[i]));
        if (i < k->d_rec.size()) { cudaFree(k->d_rec[i]); cudaFreeHost(k->h_rec[i]); }
    }
    k->d_rec.assign(nloc, nullptr);
    k->h_rec.assign(nloc, nullptr);
    k->rec_cap = 0;
    for (size_t i = 0; i < nloc; ++i) {
        PKC_CUDA(cudaSetDevice(k->comm->dev[i]));
        PKC_CUDA(cudaMalloc(&k->d_rec[i], (size_t)cap * sizeof(pk_frame_rec)));
        PKC_CUDA(cudaMallocHost(&k->h_rec[i], (size_t)cap * sizeof(pk_frame_rec)));
    }
    k->rec_cap = cap;
    return PK_OK;
}

// One SNR point of fun() (dataForPlot.cpp:43-95) with its frames sharded over the ranks of the communicator.
//   e <= 0 (exactly p frames, BASELINE configs[1]): rank r decodes a contiguous share of the global frame range; the
//          counters are combined by ONE all-reduce (pk_allreduce_point).
//   e > 0: the stop rule `count < p && countErr < e` in GLOBAL frame order: rounds of world x chunk frames, rank r takes
//          chunk r of every round; after a round the frame-error counts of the chunks are exchanged, the rank holding the
//          e-th error cuts its chunk right after it, later chunks are dropped.  The totals equal pk_kaneko_run_point on
//          one device (frames are identified by their global index = Philox counter), whatever the number of GPUs.
// Every rank / the one process gets the same *out.
int pk_comm_run_point(pk_comm_kaneko *k, double ebn0_db, int snr_index, uint64_t seed, long p, long e, pk_point_result *out) {
    if (!k || !out || p <= 0) return pk_set_error(PK_ERR_ARG, "Invalid values of arguments");
    pk_comm *c = k->comm;
    const int W = c->world, nloc = (int)c->dev.size();
    std::memset(out, 0, sizeof(*out));
    int rc;
    for (int i = 0; i < nloc; ++i) {
        PKC_CUDA(cudaSetDevice(c->dev[i]));
        PKC_CUDA(cudaMemsetAsync(k->d_tot[i], 0, sizeof(pk_point_result), c->streams[i]));
    }
    if (e <= 0) {
        const long per = (p + W - 1) / W;
        for (int i = 0; i < nloc; ++i) {
            const long r = c->rank0 + i, first = r * per, cnt = std::max(0L, std::min(per, p - first));
            if (cnt > 0 && (rc = pk_kaneko_run_frames_dev(k->decs[i], ebn0_db, snr_index, seed, (uint64_t)first, cnt, nullptr,
                                                          reinterpret_cast<uint64_t *>(k->d_tot[i]), c->streams[i])))
                return rc;
        }
        if ((rc = pk_allreduce_point(c, reinterpret_cast<pk_point_result *const *>(k->d_tot.data())))) return rc;
        PKC_CUDA(cudaSetDevice(c->dev[0]));
        PKC_CUDA(cudaMemcpyAsync(out, k->d_tot[0], sizeof(*out), cudaMemcpyDeviceToHost, c->streams[0]));
        return pk_comm_sync(c);
    }
    // finite e: rounds, records scanned in global frame order
    long done = 0, chunk = 4096;
    unsigned long long errs = 0;
    const int n = k->n;
    std::vector<unsigned long long> cnt_err((size_t)W);
    pk_point_result loc;
    std::memset(&loc, 0, sizeof loc);
    auto add = [&](const pk_frame_rec &r) {
        const uint64_t run = (uint64_t)r.trials - ((r.flags & PK_FLAG_EARLY_RETURN) ? 1 : 0);
        loc.frames += 1;
        loc.frame_errors += (r.flags & PK_FLAG_FRAME_ERROR) ? 1 : 0;
        loc.bit_errors += r.bit_errors;
        loc.trials += r.trials;
        loc.cmp += run * (uint64_t)(n + 6) + r.extra_cmp;
        loc.sum += run * (uint64_t)(n + 1) + r.extra_sum;
        loc.max_trials_seen = std::max<uint64_t>(loc.max_trials_seen, r.trials);
        loc.flags_or |= r.flags;
    };
    while (done < p && errs < (unsigned long long)e) {
        const long span = std::min((long)W * chunk, p - done);
        if ((rc = ensure_recs(k, chunk))) return rc;
        std::vector<long> cnt(nloc, 0);
        for (int i = 0; i < nloc; ++i) {
            const long r = c->rank0 + i, first = done + std::min(r * chunk, span);
            cnt[i] = std::max(0L, std::min(chunk, span - r * chunk));
            if (cnt[i] <= 0) continue;
            PKC_CUDA(cudaSetDevice(c->dev[i]));
            if ((rc = pk_kaneko_run_frames_dev(k->decs[i], ebn0_db, snr_index, seed, (uint64_t)first, cnt[i], k->d_rec[i], nullptr, c->streams[i])))
                return rc;
            PKC_CUDA(cudaMemcpyAsync(k->h_rec[i], k->d_rec[i], (size_t)cnt[i] * sizeof(pk_frame_rec), cudaMemcpyDeviceToHost, c->streams[i]));
        }
        if ((rc = pk_comm_sync(c))) return rc;
        std::fill(cnt_err.begin(), cnt_err.end(), 0ull);
        for (int i = 0; i < nloc; ++i)
            for (long f = 0; f < cnt[i]; ++f) cnt_err[(size_t)(c->rank0 + i)] += (k->h_rec[i][f].flags & PK_FLAG_FRAME_ERROR) ? 1 : 0;
        if ((rc = allreduce_host_vec(c, cnt_err))) return rc;
        unsigned long long before = errs;
        for (int r = 0; r < W; ++r) {
            const int i = r - c->rank0;
            if (i >= 0 && i < nloc && before < (unsigned long long)e) {
                // frames of this chunk count until the e-th error of the point (inclusive)
                for (long f = 0; f < cnt[i]; ++f) {
                    add(k->h_rec[i][f]);
                    if ((k->h_rec[i][f].flags & PK_FLAG_FRAME_ERROR) && before + 1 >= (unsigned long long)e) { before = (unsigned long long)e; break; }
                    before += (k->h_rec[i][f].flags & PK_FLAG_FRAME_ERROR) ? 1 : 0;
                }
                if (before >= (unsigned long long)e) break;
            } else {
                before += cnt_err[(size_t)r];
                if (before >= (unsigned long long)e) break;
            }
        }
        for (int r = 0; r < W; ++r) errs += cnt_err[(size_t)r];
        done += span;
        chunk = std::min(chunk * 2, 1L << 16);
    }
    if (c->dev.size() == 1) {
        // rank mode: this process scanned its own chunks only -- partial totals -> device -> all-reduce
        PKC_CUDA(cudaSetDevice(c->dev[0]));
        PKC_CUDA(cudaMemcpyAsync(k->d_tot[0], &loc, sizeof(loc), cudaMemcpyHostToDevice, c->streams[0]));
        if ((rc = pk_allreduce_point(c, reinterpret_cast<pk_point_result *const *>(k->d_tot.data())))) return rc;
        PKC_CUDA(cudaMemcpyAsync(out, k->d_tot[0], sizeof(*out), cudaMemcpyDeviceToHost, c->streams[0]));
        return pk_comm_sync(c);
    }
    *out = loc;   // one process holds every chunk: its scan already is the global total
    return pk_comm_sync(c);
}

}  // extern "C"
