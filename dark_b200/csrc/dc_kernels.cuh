// dc_kernels.cuh — distance coding + move-to-front of a BWT block on the GPU.
//
// Replaces `bwt::dc::encode(&output, suf, &mut self.mtf)` and the (distance, Context) items its iterator yields
// (/root/reference/src/block/dc.rs:52, 82-85).  The code being replaced is third-party (`compress::bwt::dc`,
// `compress::bwt::mtf`, Cargo.toml:18) and absent from the reference tree: PARITY UNPINNED — the kernels are checked
// against the test suite's CPU restatement of upstream rust-compress (written from memory) and by encode -> decode round trips.
//
// What is computed (n bytes of BWT in, all positions 0-based):
//   distances[i] = n (filler) unless i is the LAST position of a run of equal bytes; then with s = bwt[i], i' = the next
//   occurrence of s (start of its next run) and rank = number of distinct symbols in (i, i'):
//       distances[i] = i' - i - rank - 1;       for the last run of s: n - i - rank - 1, rank = final MTF rank of s
//   rank is exactly what mtf.encode(s) returns at i'.  init[s] = first occurrence of s (n if absent).
//   The item stream has one entry per run, in order: (run end, distance, symbol, last_rank = the MTF rank at which the
//   run began, 0 for a symbol's first run); distance_limit = n - run end.
//
// The sequential MTF list is replaced by "last occurrence" tables: the rank of s at position i' is the number of
// symbols whose last occurrence before i' is later than that of s.
//   k_dc_tables   per 4,096-byte block: last occurrence of every symbol inside the block, first occurrences, run count;
//                 fills distances[] with the filler                                   (n read, 4n written)
//   k_dc_scan_*   per symbol: running "last occurrence before block b" over the blocks, in two levels  (n/16 bytes)
//   k_dc_ranks    one warp per block walks its run starts with the last-occurrence table in registers (alphabets of up to
//                 32 symbols: one shuffle, one compare, one ballot per run) or in shared memory (8 entries per lane, one
//                 compare each, one warp reduction): rank, distance of the previous run of s, run record
//   k_dc_final    last runs and the final MTF order;   k_dc_stream   run ends + their distances, compacted
#pragma once

#include "common.cuh"

namespace dark {

constexpr int kDcBlock = 4096;

struct DcInfoDev {  // mirrors the head of dark_bwt_dc_info
    unsigned long long init[256];
    u8 mtf_symbols[256];
    u32 num_unique;
    u32 reserved_;
    unsigned long long num_items;
};

// first[256] must be preset to 0xFFFFFFFF, *total_runs to 0.
__global__ void __launch_bounds__(256)
k_dc_tables(const u8* __restrict__ bwt, u32 n, u32* __restrict__ tab, u32* __restrict__ first, u32* __restrict__ run_counts,
            u32* __restrict__ dist, unsigned long long* __restrict__ total_runs) {
    __shared__ u32 s_last[256], s_first[256];
    __shared__ u32 s_runs;
    const int tid = threadIdx.x;
    s_last[tid] = 0u;
    s_first[tid] = 0xFFFFFFFFu;
    if (tid == 0) s_runs = 0u;
    __syncthreads();
    const u64 base = (u64)blockIdx.x * kDcBlock + (u64)tid * 16;
    u32 runs = 0;
    if (base < n) {
        u32 prev = base > 0 ? (u32)bwt[base - 1] : 0x100u;
        u32 last_pos = 0;
#pragma unroll 4
        for (int k = 0; k < 16; ++k) {
            const u64 pos = base + k;
            if (pos < n) {
                const u32 s = bwt[pos];
                if (s != prev) {
                    ++runs;
                    atomicMin(&s_first[s], (u32)pos);
                    if (k > 0) atomicMax(&s_last[prev], (u32)pos);  // the run of prev ended at pos-1 (stored +1)
                }
                prev = s;
                last_pos = (u32)pos;
                dist[pos] = n;
            }
        }
        atomicMax(&s_last[prev], last_pos + 1u);  // this thread's last byte (its run may go on in the next thread's bytes)
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) runs += __shfl_xor_sync(0xffffffffu, runs, o);
    if ((tid & 31) == 0 && runs) atomicAdd(&s_runs, runs);
    __syncthreads();
    tab[(size_t)blockIdx.x * 256 + tid] = s_last[tid];
    if (s_first[tid] != 0xFFFFFFFFu) atomicMin(&first[tid], s_first[tid]);
    if (tid == 0) {
        run_counts[blockIdx.x] = s_runs;
        atomicAdd(total_runs, (unsigned long long)s_runs);
    }
}

// tab[b][c] (last occurrence + 1 of c inside block b, 0 = none)  ->  last occurrence + 1 of c BEFORE block b, in two
// levels so that no thread walks more than kDcGroup + ngroups rows (one CTA walking all n / 4096 rows took 5 ms of a
// 16 ms stage on a 256 MiB block):
//   k_dc_scan_groups  CTA g, thread c: running last occurrence over the kDcGroup blocks of group g, starting from "none";
//                     group_last[g][c] = the group's own last occurrence
//   k_dc_scan_carry   one CTA: group_last[g][c] -> last occurrence before group g; final_last[c] = over the whole input
// A reader takes tab[b][c], or the carry of b's group where the group has not seen c before b.
constexpr int kDcGroup = 256;
__global__ void __launch_bounds__(256) k_dc_scan_groups(u32* __restrict__ tab, u32 nblocks, u32* __restrict__ group_last) {
    const int c = threadIdx.x;
    const u32 b0 = blockIdx.x * kDcGroup, b1 = min(nblocks, b0 + kDcGroup);
    u32 run = 0;
    u32 b = b0;
    for (; b + 8 <= b1; b += 8) {
        u32 t[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) t[k] = tab[(size_t)(b + k) * 256 + c];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            tab[(size_t)(b + k) * 256 + c] = run;
            if (t[k]) run = t[k];
        }
    }
    for (; b < b1; ++b) {
        const u32 t = tab[(size_t)b * 256 + c];
        tab[(size_t)b * 256 + c] = run;
        if (t) run = t;
    }
    group_last[(size_t)blockIdx.x * 256 + c] = run;
}
__global__ void __launch_bounds__(256) k_dc_scan_carry(u32* __restrict__ group_last, u32 ngroups, u32* __restrict__ final_last) {
    const int c = threadIdx.x;
    u32 run = 0;
    for (u32 g = 0; g < ngroups; ++g) {
        const u32 t = group_last[(size_t)g * 256 + c];
        group_last[(size_t)g * 256 + c] = run;
        if (t) run = t;
    }
    final_last[c] = run;
}

// Dense codes of the symbols that occur (first[c] != ~0): alpha[0..255] = byte -> code, alpha[256..511] = code -> byte,
// alpha[512] = sigma.  One CTA of 256 threads.
__global__ void __launch_bounds__(256) k_dc_alphabet(const u32* __restrict__ first, u32* __restrict__ alpha) {
    __shared__ u32 s_warp[8];
    const int c = threadIdx.x, lane = c & 31, warp = c >> 5;
    const bool present = first[c] != 0xFFFFFFFFu;
    const u32 bal = __ballot_sync(0xffffffffu, present);
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    u32 base = 0;
    for (int w = 0; w < warp; ++w) base += s_warp[w];
    const u32 code = base + __popc(bal & ((1u << lane) - 1u));
    alpha[c] = present ? code : 0xFFu;
    if (present) alpha[256 + code] = (u32)c;
    if (c == 255) alpha[512] = base + __popc(bal);
}

// One warp per block.  run_offsets = exclusive scan of the blocks' run counts.
// Alphabets of up to 32 symbols (DNA, most text) keep the table in registers, one symbol per lane: the rank of a run start
// is one shuffle, one compare and one ballot.  Larger alphabets keep 256 entries in shared memory, 8 per lane.
__global__ void __launch_bounds__(256)
k_dc_ranks(const u8* __restrict__ bwt, u32 n, const u32* __restrict__ tab, const u32* __restrict__ group_carry,
           const u32* __restrict__ run_offsets, u32 nblocks, const u32* __restrict__ alpha, u32* __restrict__ dist, u32* __restrict__ run_start, u8* __restrict__ run_sym,
           u8* __restrict__ run_rank) {
    __shared__ u32 s_tab[8][256];
    __shared__ u8 s_dense[256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    s_dense[threadIdx.x] = (u8)alpha[threadIdx.x];
    const u32 sigma = alpha[512];
    __syncthreads();
    const u32 block = blockIdx.x * 8 + warp;
    if (block >= nblocks) return;
    const bool small = sigma <= 32u;
    u32* last = s_tab[warp];
    u32 lastreg = 0;  // small alphabets: last occurrence + 1 of the symbol with dense code `lane`
    const u32* carry_row = group_carry + (size_t)(block / kDcGroup) * 256;  // where the block's group had not met the symbol yet
    if (small) {
        if ((u32)lane < sigma) {
            const u32 c = alpha[256 + lane];
            const u32 t = tab[(size_t)block * 256 + c];
            lastreg = t ? t : carry_row[c];
        }
    } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const u32 t = tab[(size_t)block * 256 + lane + 32 * k];
            last[lane + 32 * k] = t ? t : carry_row[lane + 32 * k];
        }
    }
    __syncwarp();
    u32 run_idx = run_offsets[block];
    const u64 block_base = (u64)block * kDcBlock;
    u32 carry = block_base > 0 ? (u32)bwt[block_base - 1] : 0x100u;
    for (int chunk = 0; chunk < kDcBlock / 32; ++chunk) {
        const u64 pos0 = block_base + (u64)chunk * 32;
        if (pos0 >= n) break;
        const u64 pos = pos0 + lane;
        const bool valid = pos < n;
        const u32 byte = valid ? (u32)bwt[pos] : 0u;
        u32 prevb = __shfl_up_sync(0xffffffffu, byte, 1);
        if (lane == 0) prevb = carry;
        u32 mask = __ballot_sync(0xffffffffu, valid && byte != prevb);
        // run records of this chunk are written together: lane q keeps the q-th run start of the chunk
        const u32 nruns = __popc(mask);
        u32 rec_start = 0, rec_sym = 0, rec_rank = 0;
        u32 q = 0;
        while (mask) {
            const int j = __ffs(mask) - 1;
            mask &= mask - 1;
            const u32 s = __shfl_sync(0xffffffffu, byte, j), ps = __shfl_sync(0xffffffffu, prevb, j);
            const u32 i = (u32)(pos0 + j);
            u32 old, rank;
            if (small) {
                if (ps < 0x100u && (u32)lane == (u32)s_dense[ps]) lastreg = i;  // the run before this one ended at i-1 (stored +1)
                old = __shfl_sync(0xffffffffu, lastreg, (int)s_dense[s]);
                rank = __popc(__ballot_sync(0xffffffffu, lastreg > old));
            } else {
                if (lane == 0 && ps < 0x100u) last[ps] = i;
                __syncwarp();
                old = last[s];  // end + 1 of the previous run of s (0: this is its first)
                u32 cnt = 0;
#pragma unroll
                for (int k = 0; k < 8; ++k) cnt += last[lane + 32 * k] > old ? 1u : 0u;
                rank = __reduce_add_sync(0xffffffffu, cnt);
                __syncwarp();
            }
            DARK_ASSERT(old <= i && (old == 0 || i >= old + rank));
            if (lane == 0 && old) dist[old - 1] = i - (old - 1u) - rank - 1u;
            if ((u32)lane == q) {
                rec_start = i;
                rec_sym = s;
                rec_rank = old ? rank : 0u;
            }
            ++q;
        }
        DARK_ASSERT((u64)run_idx + nruns <= (u64)n);
        if ((u32)lane < nruns) {
            run_start[run_idx + lane] = rec_start;
            run_sym[run_idx + lane] = (u8)rec_sym;
            run_rank[run_idx + lane] = (u8)rec_rank;
        }
        run_idx += nruns;
        carry = __shfl_sync(0xffffffffu, byte, 31);
    }
}

// Last runs + final MTF order + init.  One CTA of 256 threads.
__global__ void __launch_bounds__(256)
k_dc_final(u32 n, const u32* __restrict__ final_last, const u32* __restrict__ first, u32* __restrict__ dist,
           const unsigned long long* __restrict__ total_runs, DcInfoDev* __restrict__ info) {
    __shared__ u32 s_last[256];
    __shared__ u32 s_present;
    const int c = threadIdx.x;
    const u32 mine = final_last[c];
    s_last[c] = mine;
    if (c == 0) s_present = 0;
    __syncthreads();
    u32 rank = 0;
    for (int k = 0; k < 256; ++k) rank += s_last[k] > mine ? 1u : 0u;
    if (mine) {
        dist[mine - 1] = n - (mine - 1u) - rank - 1u;
        info->mtf_symbols[rank] = (u8)c;
        atomicAdd(&s_present, 1u);
    }
    info->init[c] = mine ? (unsigned long long)first[c] : (unsigned long long)n;
    __syncthreads();
    if (!mine) {  // absent symbols fill the tail of the list in ascending order (any order will do: they are never looked up)
        u32 below = 0;
        for (int k = 0; k < c; ++k) below += s_last[k] == 0 ? 1u : 0u;
        info->mtf_symbols[s_present + below] = (u8)c;
    }
    if (c == 0) {
        info->num_unique = s_present;
        info->reserved_ = 0;
        info->num_items = *total_runs;
    }
}

// The item stream: run r ends just before run r+1 starts.
__global__ void __launch_bounds__(256)
k_dc_stream(const u32* __restrict__ run_start, const unsigned long long* __restrict__ total_runs, u32 n, const u32* __restrict__ dist,
            u32* __restrict__ out_pos, u32* __restrict__ out_dist) {
    const u64 R = *total_runs;
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x; r < R; r += stride) {
        const u32 end = (r + 1 < R ? run_start[r + 1] : n) - 1u;
        out_pos[r] = end;
        out_dist[r] = dist[end];
    }
}

}  // namespace dark
