/*
 * refine_split.inl -- the body of one Clustering::split (Preprocessor.cpp:590-684) as executed by one CTA: the two
 * weightedSample draws, the direction, the projections, the sort, both prefix-variance sweeps and the first-minimum argmin.
 * Textually included by the refinement kernels of refine.cuh (k_refine: one CTA owns a Clustering object; k_refine_mt: CTAs
 * pull clusters of any object from a queue), which provide the names used here:
 *   sm (RfShared), tid / lane / warp, I (RfInst *), nr / nrP / nq / tS / TV, lw, Vi, icw,
 *   begin, n, small, list (the cluster's VRLs, read) and listDst (where the sorted order goes; may alias list),
 *   Xs / Xd (the cluster's columns in list order, read / in sorted order, written), keys, wA / WfA / WrA, pairsF / pairsR,
 *   srcG / posTmp, scanIt, ringPhase, tPhase, and sm.u1 / sm.u2 / sm.nodeKey / sm.nodePos (the cluster's sample sub-stream).
 * Gangs (k_refine_mt only): a large cluster is split by gG CTAs together (gMi = this CTA's index in the gang, 0 = leader;
 * gG == 1: no gang).  Every member runs the cheap sequential parts redundantly (the two draws, the direction: pure functions
 * of the cluster), the column loops of the projections, the sort passes, the sorted copy and the step ranges of the variance
 * sweeps are dealt to the members, the scratch arrays are the leader's, and RF_GANG_SYNC() -- a counter in the node record --
 * separates the phases.  The sweeps of a member start from carries S_r(k0 - 1) that a cheap first pass sums per step range
 * (the same re-association of the double sums as the segment carries of the batched pipeline, clustering.cu).
 * A `break` leaves the enclosing loop with sm.err set.  On exit, warp w's candidate is in sm.rb / sm.rs / sm.ri [w], the
 * sorted (projection, vrl) keys are in keys[], and pairsF / pairsR hold the head / tail variance pairs.
 */
#ifndef RF_GANG_SYNC
/* all members of the gang have finished the phase, and what they wrote is visible (the fence invalidates this SM's L1) */
#define RF_GANG_SYNC()                                                                              \
    do {                                                                                            \
        __syncthreads();                                                                            \
        if (gG > 1u) {                                                                              \
            gPhase += gG;                                                                           \
            if (tid == 0) {                                                                         \
                const long long gw0_ = clock64();                                                   \
                __threadfence();                                                                    \
                atomicAdd(gCnt, 1u);                                                                \
                while (*(volatile uint32_t *) gCnt < gPhase) __nanosleep(64);                       \
                __threadfence();                                                                    \
                sm.mtClk[4] += (unsigned long long) (clock64() - gw0_); sm.mtClk[5]++;              \
            }                                                                                       \
            __syncthreads();                                                                        \
        }                                                                                           \
    } while (0)
#endif
            /* ---- weightedSample x 2 (597-602, 1534-1580) ---- */
            const uint32_t numChunks = (n + RF_CHUNK - 1) / RF_CHUNK;
            uint32_t idx[2] = {0, 0};
            {
                uint32_t cT = 0;                                    /* chunk whose running sums acc[] holds */
                for (int draw = 0; draw < 2; draw++) {
                    /* running sums: chunkEnd[c] = sum after chunk c.  Draw 1 resumes in the chunk of the first centre. */
                    uint32_t cFrom = 0;
                    if (draw == 1) {
                        const uint32_t l1 = idx[0] - cT * RF_CHUNK, cnt = min((uint32_t) RF_CHUNK, n - cT * RF_CHUNK);
                        if (tid == 0) {
                            float a = l1 ? sm.acc[l1 - 1] : (cT ? sm.chunkEnd[cT - 1] : 0.0f);
                            a += 0.0f; sm.acc[l1] = a;
                            for (uint32_t i = l1 + 1; i < cnt; i++) { a += sm.sw[0][i]; sm.acc[i] = a; }
                            sm.chunkEnd[cT] = a;
                        }
                        cFrom = cT + 1;
                    }
                    if (draw == 0 && numChunks == 1) {
                        if (tid < n) sm.sw[0][tid] = icw[list[tid]];
                        if (tid + RF_THREADS < n) sm.sw[0][tid + RF_THREADS] = icw[list[tid + RF_THREADS]];
                        __syncthreads();
                        if (tid == 0) { float a = 0.0f; for (uint32_t i = 0; i < n; i++) { a += sm.sw[0][i]; sm.acc[i] = a; } sm.chunkEnd[0] = a; }
                    } else if (cFrom < numChunks) {                 /* double-buffered gather / chain over the remaining chunks */
                        __syncthreads();
                        for (uint32_t i = tid; i < min((uint32_t) RF_CHUNK, n - cFrom * RF_CHUNK); i += RF_THREADS) sm.sw[1][i] = icw[list[cFrom * RF_CHUNK + i]];
                        for (uint32_t c = cFrom; c < numChunks; c++) {
                            const uint32_t b = (c - cFrom + 1) & 1;
                            __syncthreads();
                            if (c + 1 < numChunks)
                                for (uint32_t i = tid; i < min((uint32_t) RF_CHUNK, n - (c + 1) * RF_CHUNK); i += RF_THREADS) sm.sw[b ^ 1][i] = icw[list[(c + 1) * RF_CHUNK + i]];
                            if (tid == 0) {
                                float a = c ? sm.chunkEnd[c - 1] : 0.0f;
                                const uint32_t cnt = min((uint32_t) RF_CHUNK, n - c * RF_CHUNK);
                                for (uint32_t i = 0; i < cnt; i++) a += sm.sw[b][i];
                                sm.chunkEnd[c] = a;
                            }
                        }
                    }
                    __syncthreads();
                    const float weightSum = sm.chunkEnd[numChunks - 1];
                    const float alpha = (draw == 0 ? sm.u1 : sm.u2) * weightSum;
                    if (tid == 0 && !(weightSum > 0)) sm.flags |= 2u;
                    uint32_t cNew = 0;
                    if (numChunks > 1) {                            /* first chunk whose end sum reaches alpha */
                        for (uint32_t c = tid; c < numChunks; c += RF_THREADS) if (sm.chunkEnd[c] >= alpha) atomicMin(&sm.found, c);
                        __syncthreads();
                        cNew = sm.found == 0xffffffffu ? 0u : sm.found;
                        __syncthreads();
                        if (tid == 0) sm.found = 0xffffffffu;
                        if (!(draw == 1 && cNew == cT)) {           /* re-chain that chunk, keeping its running sums */
                            const uint32_t cnt = min((uint32_t) RF_CHUNK, n - cNew * RF_CHUNK);
                            for (uint32_t i = tid; i < cnt; i += RF_THREADS) sm.sw[0][i] = icw[list[cNew * RF_CHUNK + i]];
                            __syncthreads();
                            if (tid == 0) { float a = cNew ? sm.chunkEnd[cNew - 1] : 0.0f; for (uint32_t i = 0; i < cnt; i++) { a += sm.sw[0][i]; sm.acc[i] = a; } }
                        }
                        __syncthreads();
                    }
                    cT = cNew;
                    {
                        const uint32_t cnt = min((uint32_t) RF_CHUNK, n - cT * RF_CHUNK);
                        for (uint32_t i = tid; i < cnt; i += RF_THREADS) if (sm.acc[i] >= alpha) atomicMin(&sm.found, i);
                    }
                    __syncthreads();
                    if (sm.found == 0xffffffffu) { idx[draw] = 0; if (tid == 0) sm.flags |= 2u; }
                    else idx[draw] = cT * RF_CHUNK + sm.found;
                    __syncthreads();
                    if (tid == 0) sm.found = 0xffffffffu;
                }
            }
            __syncthreads();
            if (sm.flags & 2u) { if (tid == 0) sm.err = RF_ERR_WEIGHTS; break; }
            RF_TICK(0);

            /* ---- direction (604-623) ----  Objects of more than RF_MAXROWS rows keep the two centre columns and the direction in
             * the tile (free until the projections stage columns into its first 64 KB; the direction sits behind that) */
            const bool bigRows = nr > (uint32_t) RF_MAXROWS;
            float *const c1p = bigRows ? sm.tile : sm.c1, *const c2p = bigRows ? sm.tile + RF_MAXROWS_BIG : sm.c2;
            float *const sdirp = bigRows ? sm.tile + 2 * RF_THREADS * 16 : sm.sdir;
            for (uint32_t r = tid; r < nr; r += RF_THREADS) { c1p[r] = Xs[(size_t) idx[0] * nrP + r]; c2p[r] = Xs[(size_t) idx[1] * nrP + r]; }
            __syncthreads();
            if (tid == 0) { float a = 0; for (uint32_t r = 0; r < nr; r++) a += fabsf(c1p[r]) * fabsf(c1p[r]); sm.norm[0] = sqrtf(a); }
            else if (tid == 32) { float a = 0; for (uint32_t r = 0; r < nr; r++) a += fabsf(c2p[r]) * fabsf(c2p[r]); sm.norm[1] = sqrtf(a); }
            else if (tid == 64) { float a = 0; for (uint32_t r = 0; r < nr; r++) { const float d = c2p[r] - c1p[r]; a += fabsf(d) * fabsf(d); } sm.norm[2] = sqrtf(a); }
            __syncthreads();
            if (sm.norm[0] != 0 && sm.norm[1] != 0 && sm.norm[2] != 0) {
                const float dl = sm.norm[2];
                for (uint32_t r = tid; r < nrP; r += RF_THREADS) sdirp[r] = r < nr ? (c2p[r] - c1p[r]) / dl : 0.0f;
            } else {
                /* degenerate centres: direction uniform on the n-sphere, warp::squareToStdNormal(next2D()).x per row (616-622);
                 * log and cos evaluated in double and rounded (pinned transcendental, same on the host path and in the oracle) */
                for (;;) {
                    const uint32_t base = sm.nodePos, nkey = sm.nodeKey;
                    for (uint32_t r = tid; r < nr; r += RF_THREADS) {
                        const float s1 = alvrl_rng_uniform(nkey, base + 2 * r), s2 = alvrl_rng_uniform(nkey, base + 2 * r + 1);
                        const float rr = sqrtf(-2 * (float) log((double) (1 - s1))), phi = (float) (2 * M_PI * s2);
                        c1p[r] = (float) cos((double) phi) * rr;
                    }
                    __syncthreads();
                    if (tid == 0) {
                        sm.nodePos = base + 2 * nr; sm.degenerate++;
                        float a = 0; for (uint32_t r = 0; r < nr; r++) a += fabsf(c1p[r]) * fabsf(c1p[r]);
                        sm.norm[2] = sqrtf(a);
                    }
                    __syncthreads();
                    if (sm.norm[2] != 0) break;
                }
                const float dl = sm.norm[2];
                for (uint32_t r = tid; r < nrP; r += RF_THREADS) sdirp[r] = r < nr ? c1p[r] / dl : 0.0f;
            }
            __syncthreads();

            RF_TICK(1);
            /* ---- projections (625-640).  The columns are staged through the tile in chunks with asynchronous 16-byte copies (the
             *      whole chunk is in flight at once), then thread = column sums sequentially in fp32 in row order out of shared
             *      memory (the zero padding of columns and direction adds exact zeros).  A local matrix that fits the tile in one
             *      chunk stays there for the variance sweep. ---- */
            const bool fits = n <= TV && !bigRows;
            if (fits) {
                for (uint32_t i = tid; i < n * nq; i += RF_THREADS) {
                    const uint32_t c = i / nq, q = i - c * nq;
                    rf_cp_async16(sm.tile + c * tS + 4 * (q ^ (c & 7u)), Xs + (size_t) c * nrP + 4 * q);
                }
                rf_cp_commit(); rf_cp_wait<0>();
                __syncthreads();
            }
            RF_TICK(2);
            auto emitKey = [&](uint32_t c, float pj) {
                const float q = pj + 0.0f;                                      /* -0.0 and +0.0 compare equal in the pair order */
                uint32_t b = __float_as_uint(q);
                b = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
                const uint32_t vid = list[c];
                keys[c] = ((unsigned long long) b << 32) | vid;
                if (small) sm.pos[c] = (uint16_t) c; else posTmp[vid] = c;
            };
            if (fits) {
                for (uint32_t c = tid; c < n; c += RF_THREADS) {
                    float s = 0, pj = 0;
                    const float *x = sm.tile + c * tS;
                    const uint32_t sw = c & 7u;
#pragma unroll 2
                    for (uint32_t q = 0; q < nq; q++) {
                        const float4 e = *reinterpret_cast<const float4 *>(x + 4 * (q ^ sw));
                        float a;
                        a = fabsf(e.x); s += a * a; a = fabsf(e.y); s += a * a; a = fabsf(e.z); s += a * a; a = fabsf(e.w); s += a * a;
                    }
                    const float len = sqrtf(s);
                    if (len != 0) {
                        const float y = 1.0f / len;
                        const bool lenOk = rf_div_len_ok(len);
#pragma unroll 2
                        for (uint32_t q = 0; q < nq; q++) {
                            const float4 e = rf_div4(*reinterpret_cast<const float4 *>(x + 4 * (q ^ sw)), len, y, lenOk);   /* e / len */
                            const float4 d = *reinterpret_cast<const float4 *>(sdirp + 4 * q);
                            pj += d.x * e.x; pj += d.y * e.y; pj += d.z * e.z; pj += d.w * e.w;
                        }
                    }
                    emitKey(c, pj);
                }
            } else {
                /* Too large for the tile: RF_THREADS columns at a time, thread = column, and the columns' rows pass through the tile
                 * in blocks of 16 rows (two buffers of RF_THREADS x 16 floats, asynchronous 16-byte copies: one warp instruction
                 * copies 64 contiguous bytes of each of 8 columns), so that global memory is read in whole sectors while the
                 * sequential fp32 sums -- same operations in the same order -- run out of shared memory.  Granule g of column c sits
                 * at slot g ^ ((c >> 1) & 3) of the column's 64 bytes: the LDS.128 of 8 consecutive threads cover 8 distinct slots. */
                float *const pbuf0 = sm.tile, *const pbuf1 = sm.tile + RF_THREADS * 16;
                const uint32_t nrb = (nq + 3u) >> 2;
                for (uint32_t c0 = gMi * RF_THREADS; c0 < n; c0 += gG * RF_THREADS) {
                    const uint32_t cc = min((uint32_t) RF_THREADS, n - c0);
                    auto stageRows = [&](uint32_t rb) {
                        float *dst = (rb & 1u) ? pbuf1 : pbuf0;
                        const uint32_t g0 = rb << 2, gcnt = min(4u, nq - g0);
                        for (uint32_t i = tid; i < cc * 4u; i += RF_THREADS) {
                            const uint32_t col = i >> 2, g = i & 3u;
                            if (g < gcnt) rf_cp_async16(dst + col * 16u + 4u * (g ^ ((col >> 1) & 3u)), Xs + (size_t) (c0 + col) * nrP + 4u * (g0 + g));
                        }
                        rf_cp_commit();
                    };
                    const uint32_t swz = (tid >> 1) & 3u;
                    float s = 0, pj = 0;
                    stageRows(0);
                    for (uint32_t rb = 0; rb < nrb; rb++) {                         /* pass 1: |column|^2, rows in order */
                        if (rb + 1 < nrb) { stageRows(rb + 1); rf_cp_wait<1>(); } else rf_cp_wait<0>();
                        __syncthreads();
                        if (tid < cc) {
                            const float *x = ((rb & 1u) ? pbuf1 : pbuf0) + tid * 16u;
                            const uint32_t gcnt = min(4u, nq - (rb << 2));
                            for (uint32_t g = 0; g < gcnt; g++) {
                                const float4 e = *reinterpret_cast<const float4 *>(x + 4u * (g ^ swz));
                                float a;
                                a = fabsf(e.x); s += a * a; a = fabsf(e.y); s += a * a; a = fabsf(e.z); s += a * a; a = fabsf(e.w); s += a * a;
                            }
                        }
                        __syncthreads();
                    }
                    const float len = sqrtf(s), y = 1.0f / len;
                    const bool lenOk = rf_div_len_ok(len);
                    stageRows(0);
                    for (uint32_t rb = 0; rb < nrb; rb++) {                         /* pass 2: the projection, rows in order */
                        if (rb + 1 < nrb) { stageRows(rb + 1); rf_cp_wait<1>(); } else rf_cp_wait<0>();
                        __syncthreads();
                        if (tid < cc && len != 0) {
                            const float *x = ((rb & 1u) ? pbuf1 : pbuf0) + tid * 16u;
                            const uint32_t g0 = rb << 2, gcnt = min(4u, nq - g0);
                            for (uint32_t g = 0; g < gcnt; g++) {
                                const float4 e = rf_div4(*reinterpret_cast<const float4 *>(x + 4u * (g ^ swz)), len, y, lenOk);   /* e / len */
                                const float4 d = *reinterpret_cast<const float4 *>(sdirp + 4u * (g0 + g));
                                pj += d.x * e.x; pj += d.y * e.y; pj += d.z * e.z; pj += d.w * e.w;
                            }
                        }
                        __syncthreads();
                    }
                    if (tid < cc) emitKey(c0 + tid, pj);
                }
            }
            /* ---- std::sort of (projection, vrl) pairs (641): bitonic network on the unique keys, always out of shared memory:
             *      up to RF_SORT_BLOCK keys in one piece (the tile is free when the keys live in global memory), more as
             *      block-local passes plus global steps for the strides that span blocks ---- */
            uint32_t m = 2; while (m < n) m <<= 1;
            for (uint32_t i = n + gMi * RF_THREADS + tid; i < m; i += gG * RF_THREADS) { keys[i] = ~0ull; if (small) sm.pos[i] = 0; }
            RF_GANG_SYNC();
            RF_TICK(3);
            if (small) {
                for (uint32_t k = 2; k <= m; k <<= 1) rf_sort_steps(sm.keys, m, 0, k, k >> 1, sm.pos);
            } else {
                unsigned long long *sk = reinterpret_cast<unsigned long long *>(sm.tile);
                const uint32_t blkLen = min(m, (uint32_t) RF_SORT_BLOCK), nblk = m / blkLen;
                for (uint32_t bq = gMi; bq < nblk; bq += gG) {                  /* every block sorted (direction by global index) */
                    const uint32_t blk = bq * blkLen;
                    for (uint32_t i = tid; i < blkLen; i += RF_THREADS) sk[i] = keys[blk + i];
                    __syncthreads();
                    for (uint32_t k = 2; k <= blkLen; k <<= 1) rf_sort_steps(sk, blkLen, blk, k, k >> 1, nullptr);
                    for (uint32_t i = tid; i < blkLen; i += RF_THREADS) keys[blk + i] = sk[i];
                    __syncthreads();
                }
                RF_GANG_SYNC();
                const uint32_t tPer = (m >> 1) / gG;                            /* compare-exchange pairs per member (powers of two) */
                for (uint32_t k = 2 * blkLen; k <= m; k <<= 1) {                /* merges across blocks */
                    for (uint32_t j = k >> 1; j >= blkLen; j >>= 1) {
                        for (uint32_t t = gMi * tPer + tid; t < (gMi + 1u) * tPer; t += RF_THREADS) {
                            const uint32_t lo = ((t & ~(j - 1)) << 1) | (t & (j - 1)), hi = lo | j;
                            const unsigned long long a = keys[lo], b = keys[hi];
                            if ((a > b) == ((lo & k) == 0)) { keys[lo] = b; keys[hi] = a; }
                        }
                        RF_GANG_SYNC();
                    }
                    for (uint32_t bq = gMi; bq < nblk; bq += gG) {
                        const uint32_t blk = bq * blkLen;
                        for (uint32_t i = tid; i < blkLen; i += RF_THREADS) sk[i] = keys[blk + i];
                        __syncthreads();
                        rf_sort_steps(sk, blkLen, blk, k, blkLen >> 1, nullptr);
                        for (uint32_t i = tid; i < blkLen; i += RF_THREADS) keys[blk + i] = sk[i];
                        __syncthreads();
                    }
                    RF_GANG_SYNC();
                }
            }
            RF_TICK(4);
            /* ---- sorted list, weights and prefix weights (forward and reverse order); the gang leader alone ---- */
            if (gMi == 0u) {
                double cW[2] = {0, 0};
                for (uint32_t k0 = 0; k0 < n; k0 += RF_THREADS) {
                    const uint32_t k = k0 + tid, cnt = min((uint32_t) RF_THREADS, n - k0);
                    double v[2] = {0, 0}, tot[2];
                    if (k < n) {
                        const uint32_t vf = (uint32_t) (keys[k] & 0xffffffffull), vr = (uint32_t) (keys[n - 1 - k] & 0xffffffffull);
                        listDst[k] = vf;
                        if (!small) srcG[k] = posTmp[vf];
                        v[0] = (double) icw[vf]; v[1] = (double) icw[vr];
                        wA[k] = v[0];
                    }
                    rf_scan<2, false>(v, sm.scan[(scanIt++) & 1], (cnt + 31) / 32, n > RF_THREADS, tot);
                    if (k < n) { WfA[k] = cW[0] + v[0]; WrA[k] = cW[1] + v[1]; }
                    cW[0] += tot[0]; cW[1] += tot[1];
                }
            }
            /* the sorted copy: from the tile when the local matrix is resident, else gathered column by column (the variance
             * sweeps then stream it with bulk copies).  In a gang the members start on it at once -- they look the source column
             * up themselves (posTmp[vrl]) -- and claim chunks of 256 columns from a cursor, so that the leader, busy with the
             * prefix weights above, only takes what is left when it gets here. */
            if (fits) {
                for (uint32_t i = tid; i < n * nq; i += RF_THREADS) {
                    const uint32_t k = i / nq, q = i - k * nq, pc = sm.pos[k];
                    *reinterpret_cast<float4 *>(Xd + (size_t) k * nrP + 4 * q) = *reinterpret_cast<const float4 *>(sm.tile + pc * tS + 4 * (q ^ (pc & 7u)));
                }
            } else {
                auto copyColumns = [&](uint32_t kFrom, uint32_t kTo, uint32_t stride) {
                    for (uint32_t k4 = kFrom + warp * 4; k4 < kTo; k4 += stride) {     /* warp = column, four columns in flight */
                        const float4 *src[4];
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            const uint32_t k = min(k4 + u, kTo - 1);
                            const uint32_t sp = small ? (uint32_t) sm.pos[k] : (gG > 1u ? posTmp[(uint32_t) (keys[k] & 0xffffffffull)] : srcG[k]);
                            src[u] = reinterpret_cast<const float4 *>(Xs + (size_t) sp * nrP);
                        }
                        for (uint32_t q = lane; q < nq; q += 32) {
                            float4 e[4];
#pragma unroll
                            for (int u = 0; u < 4; u++) e[u] = src[u][q];
#pragma unroll
                            for (int u = 0; u < 4; u++) if (k4 + u < kTo) reinterpret_cast<float4 *>(Xd + (size_t) (k4 + u) * nrP)[q] = e[u];
                        }
                    }
                };
                if (gG > 1u) {
                    for (;;) {
                        __syncthreads();
                        if (tid == 0) sm.found = atomicAdd(gCursor, 1u);
                        __syncthreads();
                        const uint32_t kFrom = sm.found * 256u;
                        if (kFrom >= n) break;
                        copyColumns(kFrom, min(n, kFrom + 256u), RF_WARPS * 4);
                    }
                } else copyColumns(0, n, RF_WARPS * 4);
                __threadfence_block();
                asm volatile("fence.proxy.async;" ::: "memory");
            }
            RF_GANG_SYNC();
            RF_TICK(5);
            /* ---- calculateClusterVariance (1058-1120): the forward sweep on threads 0..255 and the reverse sweep on threads
             *      256..511, thread = row, sequential over the sorted steps (S_r is a running sum, accesses are contiguous across
             *      rows).  A local matrix that is not resident streams from the sorted copy through a two-stage ring in the tile,
             *      one bulk copy (TMA) per KC steps, the next chunk in flight while the current one is computed.  The per-step sums
             *      over rows B_k = sum_r (w_k S_r(k-1) - W_{k-1} x_r(k))^2 are reduced 16 steps at a time by a transposing
             *      shuffle reduction and across warps through a double-buffered stage ---- */
            {
                const uint32_t half = tid >> 8, hr = tid & 255u, hw = hr >> 5;
                double *Bh = reinterpret_cast<double *>(half ? pairsR : pairsF);       /* B_k lives where pairs[k] goes afterwards */
                const double *WA = half ? WrA : WfA;
                double (*part)[32][8] = reinterpret_cast<double (*)[32][8]>(&sm.sw[0][0]) + half * 2;   /* [buffer][step][warp] */
                /* steps per stage (>= 12 for nr <= 512); the step loop runs in groups of 16, so 17..31 steps would pay a
                 * second, mostly empty group per stage */
                /* more than RF_MAXROWS rows: the sweeps run once per block of 512 rows (B_k accumulates over the passes); a stage
                 * then holds the row block's slice of each of its columns, one bulk copy per column */
                const uint32_t colStride = bigRows ? 512u : nrP;
                uint32_t KC = min(32u, (uint32_t) (RF_TILE_FLOATS / 4) / colStride);
                if (KC > 16u && KC < 32u) KC = 16u;
                float *ring = sm.tile + half * (RF_TILE_FLOATS / 2);
                /* the steps of this gang member: [kBeg, kEnd) of the forward order and of the reverse order; the boundaries are
                 * symmetric (b_g + b_{G-g} = n) */
                auto gBound = [&](uint32_t g) -> uint32_t {
                    return 2u * g <= gG ? (uint32_t) (((uint64_t) g * n) / gG) : n - (uint32_t) (((uint64_t) (gG - g) * n) / gG);
                };
                const uint32_t kBeg = gG > 1u ? gBound(gMi) : 0u, kEnd = gG > 1u ? gBound(gMi + 1u) : n;
                const uint32_t nch = (kEnd - kBeg + KC - 1) / KC;
                uint32_t rb0 = 0, rbFloats = nrP;                                       /* first row and padded row count of the current row block */
                auto issue = [&](uint32_t c) {                                          /* the columns of chunk c -> ring stage c & 1 */
                    const uint32_t k0 = kBeg + c * KC, cnt = min(KC, kEnd - k0);
                    if (hr == 0) {
                        asm volatile("fence.proxy.async;" ::: "memory");
                        const uint32_t bytes = cnt * rbFloats * (uint32_t) sizeof(float);
                        rf_mbar_expect_tx(&sm.mbar[half][c & 1u], bytes);
                        const float *src = Xd + (size_t) (half ? n - k0 - cnt : k0) * nrP + rb0;
                        float *dst = ring + (c & 1u) * KC * colStride;
                        if (!bigRows) rf_bulk_load(dst, src, bytes, &sm.mbar[half][c & 1u]);
                        else for (uint32_t j = 0; j < cnt; j++) rf_bulk_load(dst + j * colStride, src + (size_t) j * nrP, rbFloats * (uint32_t) sizeof(float), &sm.mbar[half][c & 1u]);
                    }
                    if (!small && hr < cnt) {                                           /* w_k, W_{k-1} live in global memory */
                        const uint32_t k = k0 + hr, sp = half ? n - 1 - k : k;
                        rf_cp_async8(&sm.stepW[half][c & 1u][hr], wA + sp);
                        if (k) rf_cp_async8(&sm.stepWp[half][c & 1u][hr], WA + k - 1); else sm.stepWp[half][c & 1u][hr] = 0.0;
                    }
                    rf_cp_commit();
                };
                /* thread = rows hr and hr + RS (nr <= 512): the rows are folded onto the fewest warps, RS = roundup(nr / 2, 32),
                 * so that both row slots of a thread carry a row -- the sweep is issue-bound, and a warp whose second slot is
                 * empty costs as many issue slots as a full one */
                for (rb0 = 0; rb0 < (bigRows ? nr : 1u); rb0 += 512u) {
                const uint32_t nrB = bigRows ? min(512u, nr - rb0) : nr;                /* rows of this pass (block-local indices below) */
                rbFloats = bigRows ? ((nrB + 3u) & ~3u) : nrP;
                const uint32_t RS = (scr.unfoldRows && nrB <= 256u) ? ((nrB + 31u) & ~31u) : min(256u, (((nrB + 1u) >> 1) + 31u) & ~31u);
                const uint32_t rA = hr, rB = hr + RS;
                const bool actA = hr < RS && rA < nrB, actB = hr < RS && rB < nrB;
                const uint32_t nw = RS >> 5;
                /* per-row locality weights (neighbour slices in L_i): B_k = sum_r w_r (..)^2 */
                const bool rowWeighted = scr.rowW != nullptr;
                const double wRowA = (rowWeighted && actA) ? scr.rowW[I->r0 + rb0 + rA] : 1.0, wRowB = (rowWeighted && actB) ? scr.rowW[I->r0 + rb0 + rB] : 1.0;
                double SA = 0, SB = 0;
                const long long gs0_ = clock64();
                if (gG > 1u) {
                    /* carries: the row sums of every member's step range (one streaming pass over its columns, rows in step
                     * order), then S_r(kBeg - 1) = the sums of the ranges before this one */
                    /* all threads stream the range: thread = (granule of 4 rows, column slot), 8 columns in flight each, the slots'
                     * partial sums meet in shared memory (the tile is free until the main pass stages into it) */
                    {
                        const uint32_t slots = RF_THREADS / nq, q = tid % nq, slot = tid / nq;
                        double *part4 = reinterpret_cast<double *>(sm.tile);                /* [slots][nrP] */
                        for (uint32_t dirc = 0; dirc < 2; dirc++) {
                            double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
                            if (slot < slots) {
                                for (uint32_t k = kBeg + slot; k < kEnd; k += 8u * slots) {
                                    float4 e[8];
#pragma unroll
                                    for (int u = 0; u < 8; u++) {
                                        const uint32_t kk = k + (uint32_t) u * slots;
                                        e[u] = make_float4(0, 0, 0, 0);
                                        if (kk < kEnd) e[u] = __ldcg(reinterpret_cast<const float4 *>(Xd + (size_t) (dirc ? n - 1 - kk : kk) * nrP) + q);
                                    }
#pragma unroll
                                    for (int u = 0; u < 8; u++) { a0 += (double) e[u].x; a1 += (double) e[u].y; a2 += (double) e[u].z; a3 += (double) e[u].w; }
                                }
                                double *o4 = part4 + (size_t) slot * nrP + 4u * q;
                                o4[0] = a0; o4[1] = a1; o4[2] = a2; o4[3] = a3;
                            }
                            __syncthreads();
                            for (uint32_t r = tid; r < nrP; r += RF_THREADS) {
                                double t = 0;
                                for (uint32_t sl = 0; sl < slots; sl++) t += part4[(size_t) sl * nrP + r];
                                gCarry[(size_t) (dirc * gG + gMi) * RF_MAXROWS + r] = t;
                            }
                            __syncthreads();
                        }
                    }
                    if (tid == 0) sm.mtClk[6] += (unsigned long long) (clock64() - gs0_);
                    RF_GANG_SYNC();
                    if (hw < nw) {
                        for (uint32_t g = 0; g + 1u <= gMi; g++) {
                            const double *theirs = gCarry + (size_t) (half * gG + g) * RF_MAXROWS;
                            if (actA) SA += __ldcg(theirs + rA);
                            if (actB) SB += __ldcg(theirs + rB);
                        }
                    }
                }
                uint32_t buf = 0;
                const long long gs1_ = clock64();
                if (!fits && nch) issue(0);
                for (uint32_t c = 0; c < nch; c++) {
                    const uint32_t k0 = kBeg + c * KC, cnt = min(KC, kEnd - k0);
                    if (!fits) {
                        if (c + 1 < nch) { issue(c + 1); rf_cp_wait<1>(); } else rf_cp_wait<0>();
                        rf_mbar_wait(&sm.mbar[half][c & 1u], (ringPhase >> (c & 1u)) & 1u);
                        ringPhase ^= 1u << (c & 1u);
                        if (!small) asm volatile("bar.sync %0, 256;" ::"r"(1 + half) : "memory");
                    }
                    if (hw < nw) {
                        const float *stage = ring + (c & 1u) * KC * colStride;
                        /* one group of 16 steps; the variants are compile-time so that the common case -- a full group -- carries no
                         * per-step predicates and each residency mode only its own addressing.  A resident column pc keeps element r
                         * at colp[r ^ ((pc & 7) << 2)]: the granule swizzle (r / 4) ^ (pc & 7) touches bits 2..4 of r only. */
                        auto group = [&](auto kFits, auto kSmall, auto kFull, const uint32_t g) {
                            constexpr bool FITS = decltype(kFits)::value, SMALLC = decltype(kSmall)::value, FULL = decltype(kFull)::value;
                            float xa[16], xb[16]; double b[16];
#pragma unroll
                            for (int u = 0; u < 16; u++) {
                                xa[u] = 0; xb[u] = 0;
                                if (FULL || g + u < cnt) {
                                    if (FITS) {
                                        const uint32_t k = k0 + g + u, pc = sm.pos[half ? n - 1 - k : k];
                                        const float *colp = sm.tile + pc * tS;
                                        const uint32_t swz = (pc & 7u) << 2;
                                        if (actA) xa[u] = colp[rA ^ swz];
                                        if (actB) xb[u] = colp[rB ^ swz];
                                    } else {
                                        const float *colp = stage + (half ? cnt - 1 - (g + u) : g + u) * colStride;
                                        if (actA) xa[u] = colp[rA];
                                        if (actB) xb[u] = colp[rB];
                                    }
                                }
                            }
#pragma unroll
                            for (int u = 0; u < 16; u++) {
                                b[u] = 0;
                                if (FULL || g + u < cnt) {
                                    const uint32_t k = k0 + g + u, sp = half ? n - 1 - k : k;
                                    const double wk = SMALLC ? wA[sp] : sm.stepW[half][c & 1u][g + u];
                                    const double Wp = SMALLC ? (k ? WA[k - 1] : 0.0) : sm.stepWp[half][c & 1u][g + u];
                                    const double xad = (double) xa[u], xbd = (double) xb[u];
                                    const double ta = wk * SA - Wp * xad, tb = wk * SB - Wp * xbd;
                                    SA += xad; SB += xbd;
                                    b[u] = rowWeighted ? wRowA * (ta * ta) + wRowB * (tb * tb) : ta * ta + tb * tb;
                                }
                            }
                            rf_reduce16(b, lane);
                            if (lane < 16 && (FULL || g + lane < cnt)) part[buf][g + lane][hw] = b[0];
                        };
                        using T_ = std::true_type; using F_ = std::false_type;
                        for (uint32_t g = 0; g < cnt; g += 16) {
                            const bool full = g + 16 <= cnt;
                            if (fits) { if (full) group(T_(), T_(), T_(), g); else group(T_(), T_(), F_(), g); }
                            else if (small) { if (full) group(F_(), T_(), T_(), g); else group(F_(), T_(), F_(), g); }
                            else { if (full) group(F_(), F_(), T_(), g); else group(F_(), F_(), F_(), g); }
                        }
                    }
                    asm volatile("bar.sync %0, 256;" ::"r"(1 + half) : "memory");
                    if (hr < cnt) {
                        double sum = 0;
                        for (uint32_t w8 = 0; w8 < nw; w8++) sum += part[buf][hr][w8];
                        Bh[k0 + hr] = rb0 ? Bh[k0 + hr] + sum : sum;
                    }
                    buf ^= 1;
                }
                if (tid == 0 && gG > 1u) sm.mtClk[7] += (unsigned long long) (clock64() - gs1_);
                asm volatile("bar.sync %0, 256;" ::"r"(1 + half) : "memory");       /* the ring is free for the next row block */
                }                                                                       /* row blocks */
            }
            RF_GANG_SYNC();
            RF_TICK(6);
            if (gMi != 0u) break;                       /* the leader alone turns the B_k into variance pairs and finds the split */
            /* prefix pairs (1098-1106), thread = step: first = lw W_k Q_k, second = lw W_k SV_k */
            for (uint32_t dir = 0; dir < 2; dir++) {
                const double *WA = dir ? WrA : WfA;
                float2 *pairs = dir ? pairsR : pairsF;
                const double *Bh = reinterpret_cast<const double *>(pairs);
                double cQ = 0, cV = 0;
                for (uint32_t k0 = 0; k0 < n; k0 += RF_THREADS) {
                    const uint32_t cnt = min((uint32_t) RF_THREADS, n - k0), k = k0 + tid;
                    const bool on = tid < cnt;
                    double v2[2] = {0, 0}, tot2[2], Wk = 1.0;
                    if (on) {
                        const uint32_t sp = dir ? n - 1 - k : k;
                        const double wk = wA[sp];
                        Wk = WA[k];
                        if (k) { const double Wp = WA[k - 1]; v2[0] = (1.0 / wk + 1.0 / Wp) * Bh[k] / (Wk * Wk); }
                        v2[1] = Vi[(uint32_t) (keys[sp] & 0xffffffffull)];
                    }
                    rf_scan<2, false>(v2, sm.scan[(scanIt++) & 1], (cnt + 31) / 32, n > RF_THREADS, tot2);
                    if (on) pairs[k] = make_float2(k == 0 ? 0.0f : (float) (lw * (Wk * (cQ + v2[0]))), (float) (lw * ((cV + v2[1]) * Wk)));
                    cQ += tot2[0]; cV += tot2[1];
                }
            }
            __syncthreads();
            RF_TICK(7);
            /* ---- first minimum of head + tail variance (664-675) ---- */
            float best = INFINITY, second = INFINITY; uint32_t bi = 0xffffffffu;
            for (uint32_t k = 1 + tid; k < n; k += RF_THREADS) {
                const float2 h = pairsF[k - 1], tl = pairsR[n - 1 - k];
                const float v = h.x + h.y + tl.x + tl.y;
                if (v < best) { second = best; best = v; bi = k; }
                else if (v < second) second = v;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float b2 = __shfl_down_sync(0xffffffffu, best, o), s2 = __shfl_down_sync(0xffffffffu, second, o);
                const uint32_t i2 = __shfl_down_sync(0xffffffffu, bi, o);
                if (b2 < best || (b2 == best && i2 < bi)) { second = fminf(s2, best); best = b2; bi = i2; }
                else second = fminf(second, b2);
            }
            if (lane == 0) { sm.rb[warp] = best; sm.rs[warp] = second; sm.ri[warp] = bi; }
            __syncthreads();
