// perm_jit.hpp -- run-time specialisation of the permutation-run sweep (host side, no CUDA here).
//
// The generic kernel k_perm (kernels.cuh) interprets the word program out of shared memory because
// a register file cannot be indexed by a run-time qubit number; it is bound by shared-memory
// wavefronts (5 per gate per 1024 amplitudes), ~5x the HBM time for the 1434-gate ripple-carry run
// of BASELINE config 2.  For a program that is applied again and again (every input of
// Node.simulate's matrix case, every step of a benchmark) the engine therefore emits the SAME
// kernel with the gate loop unrolled into straight-line register code -- one LOP3 per gate for 32
// amplitudes -- compiles it with NVRTC for sm_100a on a background thread, and switches to it when
// ready.  Prologue (bit-sliced index words) and epilogue (register bit-matrix transpose, coalesced
// 16-byte gather/scatter) are textually the ones of k_perm.
#pragma once

#include <dlfcn.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "planner.hpp"

namespace b2svk {

inline std::string perm_jit_source(const std::vector<PGate>& exec_order, int n, int dtype) {
    std::string s;
    s.reserve(exec_order.size() * 40 + 8192);
    char buf[256];
    s += dtype == B2SV_C64 ? "typedef float2 amp_t;\n" : "typedef double2 amp_t;\n";
    s += "extern \"C\" __global__ void __launch_bounds__(256) b2sv_perm_jit(const amp_t* __restrict__ in, amp_t* __restrict__ out, unsigned long long win_mask, unsigned long long win_value) {\n";
    s += "  const unsigned tid = threadIdx.x, lane = tid & 31u;\n";
    s += "  const unsigned long long wg = (unsigned long long)blockIdx.x * 8ull + (tid >> 5);\n";
    s += "  const unsigned long long warp_base = wg << 10;\n";
    static const char* pat[5] = {"0xAAAAAAAAu", "0xCCCCCCCCu", "0xF0F0F0F0u", "0xFF00FF00u", "0xFFFF0000u"};
    for (int q = 0; q < n; ++q) {
        if (q < 5) snprintf(buf, sizeof(buf), "  unsigned w%d = ((lane >> %d) & 1u) ? 0xFFFFFFFFu : 0u;\n", q, q);
        else if (q < 10) snprintf(buf, sizeof(buf), "  unsigned w%d = %s;\n", q, pat[q - 5]);
        else snprintf(buf, sizeof(buf), "  unsigned w%d = ((wg >> %d) & 1ull) ? 0xFFFFFFFFu : 0u;\n", q, q - 10);
        s += buf;
    }
    // the sweep gathers out[j] = in[f^-1(j)]: apply the self-inverse gates in reverse order
    for (size_t gi = exec_order.size(); gi-- > 0;) {
        const PGate& g = exec_order[gi];
        std::string m;
        for (int q = 0; q < 64; ++q)
            if ((g.ctrl >> q) & 1) {
                snprintf(buf, sizeof(buf), "%sw%d", m.empty() ? "" : " & ", q);
                m += buf;
            }
        if (g.op == B2SV_OP_X) {
            if (m.empty()) snprintf(buf, sizeof(buf), "  w%d = ~w%d;\n", (int)g.t0, (int)g.t0);
            else snprintf(buf, sizeof(buf), "  w%d ^= %s;\n", (int)g.t0, m.c_str());
            s += buf;
        } else {
            if (m.empty()) snprintf(buf, sizeof(buf), "  { const unsigned d = w%d ^ w%d; w%d ^= d; w%d ^= d; }\n", (int)g.t0, (int)g.t1, (int)g.t0, (int)g.t1);
            else
                snprintf(buf, sizeof(buf), "  { const unsigned d = (w%d ^ w%d) & %s; w%d ^= d; w%d ^= d; }\n", (int)g.t0, (int)g.t1,
                         m.c_str(), (int)g.t0, (int)g.t1);
            s += buf;
        }
    }
    s += "  unsigned a[32] = {";
    for (int q = 0; q < 32; ++q) {
        if (q < n) snprintf(buf, sizeof(buf), "w%d%s", q, q == 31 ? "" : ", ");
        else snprintf(buf, sizeof(buf), "0u%s", q == 31 ? "" : ", ");
        s += buf;
    }
    s += "};\n";
    s += "  unsigned m = 0x0000FFFFu;\n"
         "  #pragma unroll\n"
         "  for (int j = 16; j != 0; j >>= 1) {\n"
         "    #pragma unroll\n"
         "    for (int k = 0; k < 32; ++k) {\n"
         "      if ((k & j) == 0) { const unsigned t = ((a[k] >> j) ^ a[k + j]) & m; a[k] ^= t << j; a[k + j] ^= t; }\n"
         "    }\n"
         "    m ^= m << (j >> 1);\n"
         "  }\n";
    s += "  #pragma unroll\n"
         "  for (int b0 = 0; b0 < 32; b0 += 8) {\n"
         "    unsigned long long src[8];\n"
         "    #pragma unroll\n"
         "    for (int u = 0; u < 8; ++u) src[u] = a[b0 + u];\n";
    for (int q = 32; q < n; ++q) {
        snprintf(buf, sizeof(buf),
                 "    #pragma unroll\n    for (int u = 0; u < 8; ++u) src[u] |= (unsigned long long)((w%d >> (b0 + u)) & 1u) << %d;\n", q, q);
        s += buf;
    }
    // live window: see k_perm (win_mask == 0: every source is read)
    s += "    amp_t v[8];\n"
         "    amp_t zero; zero.x = 0; zero.y = 0;\n"
         "    #pragma unroll\n"
         "    for (int u = 0; u < 8; ++u) v[u] = ((src[u] & win_mask) == win_value) ? in[src[u]] : zero;\n"
         "    #pragma unroll\n"
         "    for (int u = 0; u < 8; ++u) out[warp_base + (unsigned long long)(b0 + u) * 32ull + lane] = v[u];\n"
         "  }\n"
         "}\n";
    return s;
}

// ---- distributed form: the same sweep over a register sharded across GPUs -----------------------------
//
// The index map of a permutation run does not care where an amplitude lives: with the statevector split by the
// top g index positions over 2^g GPUs of one NVSwitch box, output j of rank r still is in[f^-1(r:j)] -- and the
// upper g bits of that source index say WHICH GPU holds it.  Every rank maps its peers' shards into its own
// address space (CUDA IPC, b2sv_ipc_attach) and this kernel gathers straight out of their HBM over NVLink:
// index arithmetic, the all-to-all and the write-out are ONE kernel, nothing is packed, staged or unpacked.
// Multi-controlled X gates whose TARGET is a global qubit (the carry chain of an Increment that runs across
// the rank boundary) therefore need no qubit swap at all -- almost every source is local -- and a real
// global<->local swap of k qubits is the same kernel with SWAP gates: one pass, (1 - 2^-k) of the shard
// crosses NVLink once.
//
// Words 0..nl-1 are the local index bits (as in b2sv_perm_jit), words nl..nl+g-1 the LOGICAL values of the
// global positions on this rank (physical rank bit ^ flip flag: an uncontrolled X on a global qubit is a rank
// relabel, unitaria_b200/distributed.py).  After the run the source rank of output b is the logical source
// bits ^ flip_in.  Per source rank the kernel gets the shard pointer, its live window (zero outside, not even
// materialised -- a rank that holds no amplitude at all has an empty window) and its pending lazy scalar.
// (struct DistPermArgs: kernels.cuh; mirrored textually in the generated source)
inline std::string perm_dist_source(const std::vector<PGate>& exec_order, int nl, int g, int dtype) {
    std::string s;
    s.reserve(exec_order.size() * 40 + 8192);
    char buf[320];
    const int n = nl + g;
    s += dtype == B2SV_C64 ? "typedef float2 amp_t;\n" : "typedef double2 amp_t;\n";
    s += "struct DistPermArgs { const amp_t* src[16]; unsigned long long wmask[16], wval[16]; double sre[16], sim[16]; amp_t* out; "
         "unsigned rank_logical; unsigned flip_in; int has_scale; unsigned block_rot; };\n";
    s += "extern \"C\" __global__ void __launch_bounds__(256) b2sv_perm_dist(const __grid_constant__ DistPermArgs a) {\n";
    s += "  const unsigned tid = threadIdx.x, lane = tid & 31u;\n";
    s += "  unsigned blk = blockIdx.x + a.block_rot; if (blk >= gridDim.x) blk -= gridDim.x;\n";
    s += "  const unsigned long long wg = ((unsigned long long)blk * blockDim.x + tid) >> 5;\n";
    s += "  const unsigned long long warp_base = wg << 10;\n";
    static const char* pat[5] = {"0xAAAAAAAAu", "0xCCCCCCCCu", "0xF0F0F0F0u", "0xFF00FF00u", "0xFFFF0000u"};
    for (int q = 0; q < n; ++q) {
        if (q >= nl) snprintf(buf, sizeof(buf), "  unsigned w%d = ((a.rank_logical >> %d) & 1u) ? 0xFFFFFFFFu : 0u;\n", q, q - nl);
        else if (q < 5) snprintf(buf, sizeof(buf), "  unsigned w%d = ((lane >> %d) & 1u) ? 0xFFFFFFFFu : 0u;\n", q, q);
        else if (q < 10) snprintf(buf, sizeof(buf), "  unsigned w%d = %s;\n", q, pat[q - 5]);
        else snprintf(buf, sizeof(buf), "  unsigned w%d = ((wg >> %d) & 1ull) ? 0xFFFFFFFFu : 0u;\n", q, q - 10);
        s += buf;
    }
    for (size_t gi = exec_order.size(); gi-- > 0;) {  // gather: the self-inverse gates in reverse order
        const PGate& gt = exec_order[gi];
        std::string m;
        for (int q = 0; q < 64; ++q)
            if ((gt.ctrl >> q) & 1) {
                snprintf(buf, sizeof(buf), "%sw%d", m.empty() ? "" : " & ", q);
                m += buf;
            }
        if (gt.op == B2SV_OP_X) {
            if (m.empty()) snprintf(buf, sizeof(buf), "  w%d = ~w%d;\n", (int)gt.t0, (int)gt.t0);
            else snprintf(buf, sizeof(buf), "  w%d ^= %s;\n", (int)gt.t0, m.c_str());
        } else if (m.empty()) {
            snprintf(buf, sizeof(buf), "  { const unsigned d = w%d ^ w%d; w%d ^= d; w%d ^= d; }\n", (int)gt.t0, (int)gt.t1, (int)gt.t0, (int)gt.t1);
        } else {
            snprintf(buf, sizeof(buf), "  { const unsigned d = (w%d ^ w%d) & %s; w%d ^= d; w%d ^= d; }\n", (int)gt.t0, (int)gt.t1, m.c_str(),
                     (int)gt.t0, (int)gt.t1);
        }
        s += buf;
    }
    s += "  unsigned t[32] = {";
    for (int q = 0; q < 32; ++q) {
        if (q < nl) snprintf(buf, sizeof(buf), "w%d%s", q, q == 31 ? "" : ", ");
        else snprintf(buf, sizeof(buf), "0u%s", q == 31 ? "" : ", ");
        s += buf;
    }
    s += "};\n";
    s += "  unsigned m = 0x0000FFFFu;\n"
         "  #pragma unroll\n"
         "  for (int j = 16; j != 0; j >>= 1) {\n"
         "    #pragma unroll\n"
         "    for (int k = 0; k < 32; ++k) {\n"
         "      if ((k & j) == 0) { const unsigned x = ((t[k] >> j) ^ t[k + j]) & m; t[k] ^= x << j; t[k + j] ^= x; }\n"
         "    }\n"
         "    m ^= m << (j >> 1);\n"
         "  }\n";
    s += "  amp_t zero; zero.x = 0; zero.y = 0;\n"
         "  #pragma unroll\n"
         "  for (int b0 = 0; b0 < 32; b0 += 8) {\n"
         "    unsigned long long src[8];\n"
         "    unsigned r[8];\n"
         "    #pragma unroll\n"
         "    for (int u = 0; u < 8; ++u) { src[u] = t[b0 + u]; r[u] = a.flip_in; }\n";
    for (int q = 32; q < nl; ++q) {
        snprintf(buf, sizeof(buf),
                 "    #pragma unroll\n    for (int u = 0; u < 8; ++u) src[u] |= (unsigned long long)((w%d >> (b0 + u)) & 1u) << %d;\n", q, q);
        s += buf;
    }
    for (int i = 0; i < g; ++i) {
        snprintf(buf, sizeof(buf), "    #pragma unroll\n    for (int u = 0; u < 8; ++u) r[u] ^= ((w%d >> (b0 + u)) & 1u) << %d;\n", nl + i, i);
        s += buf;
    }
    s += "    amp_t v[8];\n"
         "    #pragma unroll\n"
         "    for (int u = 0; u < 8; ++u) v[u] = ((src[u] & a.wmask[r[u]]) == a.wval[r[u]]) ? a.src[r[u]][src[u]] : zero;\n"
         "    if (a.has_scale) {\n"
         "      #pragma unroll\n"
         "      for (int u = 0; u < 8; ++u) {\n"
         "        const double sr = a.sre[r[u]], si = a.sim[r[u]], x = (double)v[u].x, y = (double)v[u].y;\n"
         "        v[u].x = x * sr - y * si; v[u].y = x * si + y * sr;\n"
         "      }\n"
         "    }\n"
         "    #pragma unroll\n"
         "    for (int u = 0; u < 8; ++u) a.out[warp_base + (unsigned long long)(b0 + u) * 32ull + lane] = v[u];\n"
         "  }\n"
         "}\n";
    return s;
}

// ---- NVRTC through dlopen: the library keeps no link-time dependency on it --------------------------

struct NvrtcApi {
    void* handle = nullptr;
    int (*CreateProgram)(void**, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
    int (*CompileProgram)(void*, int, const char* const*) = nullptr;
    int (*GetCUBINSize)(void*, size_t*) = nullptr;
    int (*GetCUBIN)(void*, char*) = nullptr;
    int (*GetProgramLogSize)(void*, size_t*) = nullptr;
    int (*GetProgramLog)(void*, char*) = nullptr;
    int (*DestroyProgram)(void**) = nullptr;
    bool ok = false;
};

inline const NvrtcApi& nvrtc_api() {
    static NvrtcApi api = [] {
        NvrtcApi a;
        const char* names[] = {"libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so",
                               "/usr/local/cuda/lib64/libnvrtc.so"};
        for (const char* nm : names) {
            a.handle = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
            if (a.handle) break;
        }
        if (!a.handle) return a;
        auto sym = [&](const char* n) { return dlsym(a.handle, n); };
        a.CreateProgram = (decltype(a.CreateProgram))sym("nvrtcCreateProgram");
        a.CompileProgram = (decltype(a.CompileProgram))sym("nvrtcCompileProgram");
        a.GetCUBINSize = (decltype(a.GetCUBINSize))sym("nvrtcGetCUBINSize");
        a.GetCUBIN = (decltype(a.GetCUBIN))sym("nvrtcGetCUBIN");
        a.GetProgramLogSize = (decltype(a.GetProgramLogSize))sym("nvrtcGetProgramLogSize");
        a.GetProgramLog = (decltype(a.GetProgramLog))sym("nvrtcGetProgramLog");
        a.DestroyProgram = (decltype(a.DestroyProgram))sym("nvrtcDestroyProgram");
        a.ok = a.CreateProgram && a.CompileProgram && a.GetCUBINSize && a.GetCUBIN && a.GetProgramLogSize &&
               a.GetProgramLog && a.DestroyProgram;
        return a;
    }();
    return api;
}

// Compile `src` for sm_100a; returns true and fills `cubin` on success, else fills `log`.
inline bool nvrtc_compile(const std::string& src, std::vector<char>& cubin, std::string& log) {
    const NvrtcApi& api = nvrtc_api();
    if (!api.ok) {
        log = "libnvrtc not available";
        return false;
    }
    void* prog = nullptr;
    if (api.CreateProgram(&prog, src.c_str(), "b2sv_perm_jit.cu", 0, nullptr, nullptr) != 0) {
        log = "nvrtcCreateProgram failed";
        return false;
    }
    const char* opts[] = {"--gpu-architecture=sm_100a", "-lineinfo"};
    const int rc = api.CompileProgram(prog, 2, opts);
    if (rc != 0) {
        size_t n = 0;
        api.GetProgramLogSize(prog, &n);
        log.resize(n);
        if (n) api.GetProgramLog(prog, &log[0]);
        api.DestroyProgram(&prog);
        return false;
    }
    size_t n = 0;
    api.GetCUBINSize(prog, &n);
    cubin.resize(n);
    const bool ok = n > 0 && api.GetCUBIN(prog, cubin.data()) == 0;
    api.DestroyProgram(&prog);
    if (const char* dump = getenv("B2SV_JIT_DUMP")) {  // debugging aid: keep the last source + cubin
        std::string base(dump);
        if (FILE* f = fopen((base + ".cu").c_str(), "w")) {
            fwrite(src.data(), 1, src.size(), f);
            fclose(f);
        }
        if (FILE* f = fopen((base + ".cubin").c_str(), "wb")) {
            fwrite(cubin.data(), 1, cubin.size(), f);
            fclose(f);
        }
    }
    if (!ok) log = "nvrtcGetCUBIN failed";
    return ok;
}

}  // namespace b2svk
