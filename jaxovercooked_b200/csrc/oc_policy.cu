// oc_policy.cu -- C ABI of the actor-critic forward (include/oc_b200.h: oc_policy_*): parameter repacking, kernel selection
// and launches.  The kernels live in oc_policy_kernels.cuh (architectures/mlp.py:11-59 as called at
// baselines/IPPO_MLP.py:435-439; scope-table row f3).
#include "oc_b200.h"

#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <new>

#include "oc_policy_kernels.cuh"

namespace {

typedef void (*pol_kernel_t)(const PParams);

// PL = 16-byte pieces per lane and row: 2 covers K <= 994, 4 K <= 2018, 8 K <= 4066 (a 12 x 13 grid)
pol_kernel_t pick_kernel(int act, int pl) {
    if (pl == 0) return act == OC_ACT_RELU ? oc_policy_kernel<OC_ACT_RELU, 0> : oc_policy_kernel<OC_ACT_TANH, 0>;
    if (act == OC_ACT_RELU) return pl == 2 ? oc_policy_kernel<OC_ACT_RELU, 2> : pl == 4 ? oc_policy_kernel<OC_ACT_RELU, 4> : oc_policy_kernel<OC_ACT_RELU, 8>;
    return pl == 2 ? oc_policy_kernel<OC_ACT_TANH, 2> : pl == 4 ? oc_policy_kernel<OC_ACT_TANH, 4> : oc_policy_kernel<OC_ACT_TANH, 8>;
}

// small batches: shorter tiles (the MMA stays M = 128, the unused rows are never read back) as long as the tiles still
// fit in one wave of CTAs
int rows_per_tile(long long M, int num_sms) {
    int rpt = kTileM;
    while (rpt > 32 && (M + rpt / 2 - 1) / (rpt / 2) <= num_sms) rpt >>= 1;
    return rpt;
}

// Launch with programmatic stream serialization unless the parameters were repacked since the last forward (then the
// set-up must not overtake the pack kernel).
int launch_policy(const oc_policy *p, pol_kernel_t kern, int grid, int threads, size_t smem, const PParams &k, cudaStream_t stream);

int gcd16(int k) {
    int g = 16;
    while (k % g) g >>= 1;
    return g;
}

}  // namespace

struct oc_policy {
    int K, A, act, part;
    int nvar, gshift, tv_stride;
    int num_sms;
    size_t smem_bytes, smem_cells;   // dynamic shared memory of the uint8-source / compact-source kernels
    __half *w1;
    uint8_t *w2img;
    PolicySmall *small;
    uint8_t *tmpl;
    int pl;
    // compact-observation source (desc->layout given): everything the policy needs from the layout is COPIED at creation
    // (record table, cell geometry), so the policy does not depend on the layout's lifetime
    bool has_cells;
    int C, cell_stride, cell_agents, n_scan, device;
    unsigned *rec;              // [OC_NUM_CELL_CODES][kRecWords]: the observation record of every cell code
    __half *emb;
    uint8_t *scan_cells;
    mutable bool repacked;      // set by create / update, cleared by the next forward
    long long shard_first, shard_global;   // oc_policy_set_shard (global == 0: off)
};

namespace {
int launch_policy(const oc_policy *p, pol_kernel_t kern, int grid, int threads, size_t smem, const PParams &k, cudaStream_t stream) {
    cudaLaunchConfig_t cfg;
    std::memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    static const bool use_pdl = [] { const char *e = std::getenv("OC_B200_POLICY_PDL"); return !(e && e[0] == '0'); }();
    cfg.numAttrs = (p->repacked || !use_pdl) ? 0 : 1;
    p->repacked = false;
    return cudaLaunchKernelEx(&cfg, kern, k) == cudaSuccess ? OC_OK : OC_ERR_CUDA;
}
}  // namespace

extern "C" {

static int policy_pack(oc_policy *p, const oc_policy_desc *d, cudaStream_t stream) {
    PackArgs a;
    for (int i = 0; i < 6; ++i) {
        if (!d->kernel[i] || !d->bias[i]) return OC_ERR_INVALID_ARG;
        a.kernel[i] = d->kernel[i];
        a.bias[i] = d->bias[i];
    }
    a.tmpl = p->tmpl; a.w1 = p->w1; a.w2img = p->w2img; a.small = p->small; a.K = p->K; a.A = p->A;
    p->repacked = true;
    oc_policy_pack_kernel<<<p->num_sms * 2, 256, 0, stream>>>(a);
    if (p->has_cells)
        oc_policy_embed_kernel<<<p->C * OC_NUM_CELL_CODES + 2, kHid2, 0, stream>>>(
            d->kernel[0], d->kernel[3], p->tmpl, p->rec, p->C, p->emb);
    return cudaGetLastError() == cudaSuccess ? OC_OK : OC_ERR_CUDA;
}

void oc_policy_destroy(oc_policy *p) {
    if (!p) return;
    cudaFree(p->w1); cudaFree(p->w2img); cudaFree(p->small); cudaFree(p->tmpl); cudaFree(p->emb); cudaFree(p->scan_cells); cudaFree(p->rec);
    delete p;
}

int oc_policy_create(const oc_policy_desc *d, oc_policy **out) {
    if (!d || !out) return OC_ERR_INVALID_ARG;
    *out = nullptr;
    if (d->obs_dim < 32 || d->obs_dim > 65535 || d->action_dim < 1 || d->action_dim > OC_POLICY_MAX_ACTIONS ||
        (d->activation != OC_ACT_TANH && d->activation != OC_ACT_RELU) ||
        (d->prng_mode != OC_PRNG_THREEFRY_LEGACY && d->prng_mode != OC_PRNG_THREEFRY_PARTITIONABLE))
        return OC_ERR_INVALID_ARG;
    int dev = 0;
    cudaDeviceProp prop;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return OC_ERR_NO_DEVICE;
    if (prop.major != 10) return OC_ERR_NO_DEVICE;
    oc_policy *p = new (std::nothrow) oc_policy();
    if (!p) return OC_ERR_INVALID_ARG;
    p->K = d->obs_dim; p->A = d->action_dim; p->act = d->activation; p->part = d->prng_mode;
    p->num_sms = prop.multiProcessorCount;
    p->device = dev;
    const int g = gcd16(p->K);
    p->gshift = g == 16 ? 4 : g == 8 ? 3 : g == 4 ? 2 : g == 2 ? 1 : 0;
    p->nvar = 16 / g;
    p->tv_stride = ((p->K + 15 + 31) / 16) * 16;     // a + K bytes of template, read in whole 16-byte pieces
    p->smem_bytes = (size_t)off_tmpl(false) + (size_t)p->nvar * p->tv_stride + 1024;
    p->smem_cells = (size_t)off_tmpl(true) + 1024;
    const int groups = (((p->K + 30) >> 4) + 31) / 32;
    p->pl = groups <= 2 ? 2 : groups <= 4 ? 4 : 8;
    if (groups > 8 || p->smem_bytes > (size_t)prop.sharedMemPerBlockOptin) { delete p; return OC_ERR_UNSUPPORTED_LAYOUT; }
    if (cudaMalloc(&p->w1, (size_t)p->K * kHid2 * sizeof(__half)) != cudaSuccess ||
        cudaMalloc(&p->w2img, 2 * kNetBytes) != cudaSuccess || cudaMalloc(&p->small, sizeof(PolicySmall)) != cudaSuccess ||
        cudaMalloc(&p->tmpl, p->K) != cudaSuccess) {
        oc_policy_destroy(p);
        return OC_ERR_CUDA;
    }
    const uint8_t *tmpl_src = d->obs_template;
    if (d->layout) {      // the env's own static image, and the (view, cell, code) embedding table for oc_policy_forward_cells
        const oc_layout *lay = d->layout;
        if (lay->host.C * OC_OBS_CHANNELS != p->K || lay->device != dev) { oc_policy_destroy(p); return OC_ERR_INVALID_ARG; }
        tmpl_src = lay->host.obs_tmpl;
        if (lay->cell_stride <= 64) {
            uint8_t sc[64] = {0};
            for (int i = 0; i < lay->host.n_scan; ++i) sc[i] = (uint8_t)(lay->host.scan[i] >> 16);
            if (cudaMalloc(&p->emb, ((size_t)lay->host.C * OC_NUM_CELL_CODES + 2) * kHid2 * sizeof(__half)) != cudaSuccess ||
                cudaMalloc(&p->scan_cells, 64) != cudaSuccess ||
                cudaMemcpy(p->scan_cells, sc, 64, cudaMemcpyHostToDevice) != cudaSuccess ||
                cudaMalloc(&p->rec, sizeof(lay->host.rec)) != cudaSuccess ||
                cudaMemcpy(p->rec, lay->host.rec, sizeof(lay->host.rec), cudaMemcpyHostToDevice) != cudaSuccess) {
                oc_policy_destroy(p);
                return OC_ERR_CUDA;
            }
            p->has_cells = true;
            p->C = lay->host.C; p->cell_stride = lay->cell_stride; p->cell_agents = lay->cell_agents; p->n_scan = lay->host.n_scan;
            if (cudaFuncSetAttribute(pick_kernel(p->act, 0), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem_cells) != cudaSuccess) {
                oc_policy_destroy(p);
                return OC_ERR_CUDA;
            }
        }
    }
    cudaError_t e = tmpl_src ? cudaMemcpy(p->tmpl, tmpl_src, p->K, cudaMemcpyHostToDevice) : cudaMemset(p->tmpl, 0, p->K);
    if (e != cudaSuccess ||
        cudaFuncSetAttribute(pick_kernel(p->act, p->pl), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem_bytes) != cudaSuccess) {
        oc_policy_destroy(p);
        return OC_ERR_CUDA;
    }
    const int rc = policy_pack(p, d, nullptr);
    if (rc != OC_OK || cudaStreamSynchronize(nullptr) != cudaSuccess) {
        oc_policy_destroy(p);
        return rc != OC_OK ? rc : OC_ERR_CUDA;
    }
    *out = p;
    return OC_OK;
}

int oc_policy_set_shard(oc_policy *p, int64_t first, int64_t global_count) {
    if (!p || first < 0 || global_count < 0 || (global_count && first >= global_count)) return OC_ERR_INVALID_ARG;
    p->shard_first = first; p->shard_global = global_count;
    return OC_OK;
}

int oc_policy_update(oc_policy *p, const oc_policy_desc *d, void *stream) {
    if (!p || !d || d->obs_dim != p->K || d->action_dim != p->A) return OC_ERR_INVALID_ARG;
    return policy_pack(p, d, static_cast<cudaStream_t>(stream));
}

int oc_policy_forward(const oc_policy *p, int64_t M, const uint8_t *obs, float *logits, float *value,
                      const uint32_t *key, int32_t *action, float *log_prob, void *stream_) {
    if (!p || M < 0 || !obs) return OC_ERR_INVALID_ARG;
    if ((action || log_prob) && !key) return OC_ERR_INVALID_ARG;
    if (M == 0) return OC_OK;
    if (M * p->A > 0x7fffffffll) return OC_ERR_INVALID_ARG;
    if ((reinterpret_cast<uintptr_t>(logits) | reinterpret_cast<uintptr_t>(value) | reinterpret_cast<uintptr_t>(key) |
         reinterpret_cast<uintptr_t>(action) | reinterpret_cast<uintptr_t>(log_prob)) & 3u)
        return OC_ERR_ALIGNMENT;
    PParams k;
    std::memset(&k, 0, sizeof(k));
    k.w1 = p->w1; k.w2img = p->w2img; k.small = p->small; k.tmpl = p->tmpl;
    const uintptr_t ob = reinterpret_cast<uintptr_t>(obs);
    k.a0 = (int)(ob & 15u);
    k.obs16 = reinterpret_cast<const uint4 *>(ob - (uintptr_t)k.a0);
    k.M = M; k.K = p->K; k.A = p->A; k.part = p->part;
    k.nvar = p->nvar; k.gshift = p->gshift; k.tv_stride = p->tv_stride;
    k.rpt = rows_per_tile(M, p->num_sms);
    k.n_tiles = (int)((M + k.rpt - 1) / k.rpt);
    k.noise_first = p->shard_global ? p->shard_first : 0; k.noise_global = p->shard_global ? p->shard_global : M;
    if (k.noise_first + M > k.noise_global || k.noise_global * p->A > 0x7fffffffll) return OC_ERR_INVALID_ARG;
    k.logits = logits; k.value = value; k.log_prob = log_prob; k.key = key; k.action = action;
    const int grid = k.n_tiles < p->num_sms ? k.n_tiles : p->num_sms;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    return launch_policy(p, pick_kernel(p->act, p->pl), grid, kObsWarps * 32, p->smem_bytes, k, stream);
}

int oc_policy_forward_cells(const oc_policy *p, int64_t N, const uint8_t *cells, float *logits, float *value,
                             const uint32_t *key, int32_t *action, float *log_prob, void *stream_) {
    if (!p || N < 0 || !cells) return OC_ERR_INVALID_ARG;
    if (!p->has_cells) return OC_ERR_UNSUPPORTED_LAYOUT;      // created without a layout, or its compact form is too wide
    if ((action || log_prob) && !key) return OC_ERR_INVALID_ARG;
    if (N == 0) return OC_OK;
    const int64_t M = 2 * N;
    if (M * p->A > 0x7fffffffll || M * p->cell_stride > 0x7fffffffll) return OC_ERR_INVALID_ARG;
    if ((reinterpret_cast<uintptr_t>(logits) | reinterpret_cast<uintptr_t>(value) | reinterpret_cast<uintptr_t>(key) |
         reinterpret_cast<uintptr_t>(action) | reinterpret_cast<uintptr_t>(log_prob)) & 3u)
        return OC_ERR_ALIGNMENT;
    PParams k;
    std::memset(&k, 0, sizeof(k));
    k.w1 = p->w1; k.w2img = p->w2img; k.small = p->small; k.tmpl = p->tmpl;
    k.M = M; k.N = N; k.K = p->K; k.A = p->A; k.part = p->part;
    k.nvar = p->nvar; k.gshift = p->gshift; k.tv_stride = p->tv_stride;
    k.noise_first = p->shard_global ? p->shard_first : 0; k.noise_global = p->shard_global ? p->shard_global : N;
    if (k.noise_first + N > k.noise_global || 2 * k.noise_global * p->A > 0x7fffffffll) return OC_ERR_INVALID_ARG;
    k.rpt = rows_per_tile(M, p->num_sms);                      // a tile holds rpt / 2 envs, both views
    k.n_tiles = (int)((N + k.rpt / 2 - 1) / (k.rpt / 2));
    k.logits = logits; k.value = value; k.log_prob = log_prob; k.key = key; k.action = action;
    k.cells = cells; k.emb = p->emb; k.scan_cells = p->scan_cells;
    k.cell_stride = p->cell_stride; k.cell_agents = p->cell_agents; k.n_scan = p->n_scan;
    k.C = p->C;
    const int grid = k.n_tiles < p->num_sms ? k.n_tiles : p->num_sms;
    return launch_policy(p, pick_kernel(p->act, 0), grid, kCellWarps * 32, p->smem_cells, k, static_cast<cudaStream_t>(stream_));
}

int oc_policy_sample(int prng_mode, int64_t M, int32_t A, const float *logits, const uint32_t *key, int32_t *action,
                     float *log_prob, void *stream) {
    if (M < 0 || A < 1 || A > OC_POLICY_MAX_ACTIONS || !logits || !key ||
        (prng_mode != OC_PRNG_THREEFRY_LEGACY && prng_mode != OC_PRNG_THREEFRY_PARTITIONABLE) || M * A > 0x7fffffffll)
        return OC_ERR_INVALID_ARG;
    if (M == 0) return OC_OK;
    oc_policy_sample_kernel<<<(unsigned)((4 * M + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        prng_mode, M, A, logits, key, action, log_prob);
    return cudaGetLastError() == cudaSuccess ? OC_OK : OC_ERR_CUDA;
}

}  // extern "C"
