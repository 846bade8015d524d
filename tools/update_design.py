"""One-off editor used in round 2 to bring DESIGN.md up to date section by section (kept for the record; running it on
an already updated file is a no-op because the section markers it looks for are gone)."""
import sys

p = "DESIGN.md"
s = open(p).read()


def replace_section(s, start_marker, end_marker, new):
    if start_marker not in s:
        return s
    a = s.index(start_marker)
    b = s.index(end_marker, a + len(start_marker))
    return s[:a] + new + s[b:]


# ------------------------------------------------------------------ section 0
s = replace_section(s, "## 0. Scope (SURVEY.md §8) and status after round 1", "## 1. The path and its boundary", '''## 0. Scope (SURVEY.md §8) and status after round 2

| §8 row | What | Status |
|---|---|---|
| a1 `MCMC` (4 copies) | K1 `sweep_replay_kernel` / `sweep_replay_int_kernel` (exact replay), K2 `msc_sweep_kernel` (production, ±J lattices), K2a `col_sweep_kernel` (production, any sparse J/h), K3 dense path (production, dense J) | **done**, parity green on B200 incl. the native sizes of all five configs |
| a2 energies | fused in K1, K4 `energy_kernel`, K4' `msc_energy_kernel` (bit-sliced) | **done**, exact on ±J, 1e-9 on real J |
| a3 LBP | K5 `lbp_kernel` with device `tanh`/`arctanh` that restate numpy's AVX-512 float64 routines operation for operation (`csrc/nlmc_npmath.h`) | **done, bit-exact**: iteration counts, divergence points, marginals and backbones equal the reference's at every λ (round 1: ≤1e-12, counts differed at marginal steps) |
| a4 `find_clusters` | host (set operations on seeds, negligible) | **done**, identical to reference incl. growth loop |
| a5 `NMC_subroutine` ×2 | host phase driver over K1 (`nmc_core.py`); production: per-site modes in K2a / K3 | **done**, bit-exact free-running |
| a6 `NMC.run` | `nlmc_b200/nmc.py` | **done**: whole runs bit-exact with K5 finding the backbones (no injected backbones, no xfail) |
| a7 `NPT.run` | `nlmc_b200/npt.py`; K6 on the device in production for every engine | **done** (replay bit-exact incl. `doNMC`; production: bit-packed, generic, hybrid and multi-GPU paths) |
| a8 `APT_preprocessor.run` | `nlmc_b200/apt_preprocessor.py` | **done** (replay + production) |
| a9 `APT_ICM.run` + clusters | K7 `icm_components_kernel` | **done** (replay + production, reference semantics incl. SURVEY D5) |
| b boundary | `include/nlmc_b200.h` (C ABI) + ctypes + the four classes | **done** |
| c oracle | `oracle/` pinned bit-for-bit to the live reference + goldens; numpy's own `tanh`/`arctanh` pinned by 9.4e6-argument digests | **done** |
| d measurement | `bench.py`, `profiles/` | **done**: roofline as a real fraction, e2e through the class API, time to target in the line, sustained leg |
| e multi-GPU | one ladder set with its β range sharded over the ranks: energies all-gathered (NCCL), identical label decisions on every rank, spins never move (`distributed.ShardedBetaLadder`); independent ladders per rank as the weak-scaling form | **done**: sharded ≡ single handle bit for bit on 2 and 8 B200 over NCCL with exchanges crossing the rank boundaries; strong + weak scaling in `bench.py` |
| P3 device-side exchange | K6 in β-label form for all three engines (`msc_label_swap_kernel`, `nlmc_exchange.cuh`) | **done**; an exchange between a bit-packed slot and an NMC row (hybrid ladder) moves 2·n bytes through the host |
| b' element-level public methods | `MCMC`, `LoopyBeliefPropagation`, `LBP_convexified`, `atanh_saturated`, `find_clusters`, `NMC_subroutine`, `MCMC_task`, `NMC_task`, `replica_energy` (`nlmc_b200/path_methods.py`) | **done** |
| f1 instance I/O | `nlmc_b200/instances.py` | **done** |
| f2 time to target | `tools/time_to_target.py`: Chimera-128 droplet, DCL C8, **Wishart N=36 α=0.50**, both arms, in the bench line | **done** |
| f3 artefacts | `.npy` files; PNGs behind an optional matplotlib import | **done** |
| f4 LBP by-products, generator | `nlmc_lbp_byproducts`, `instances.contrived_wishart_tree` | **done**; the `find_clusters` growth loop stays on the host (it never runs with any shipped parameter set) |

Open items are listed in §9.

''')

# ------------------------------------------------------------------ LBP parity
s = replace_section(s, '### LBP parity (the one place where "bit-exact" is not well defined)', "## 3. Data layout in HBM", '''### LBP parity (numpy's `tanh` / `arctanh` restated on the device)

`LBP_convexified` is called with `tolerance = np.finfo(float).eps` (README, `nmc.py:445`): the loop stops only at an exact
floating-point fixed point, and "diverged" (`iteration == max_iterations-1`, `nmc.py:142-146`) ends the λ schedule.  Whether a
marginally convergent λ step reaches a fixed point or a 1-ulp limit cycle depends on the last bit of `tanh`/`arctanh`
(glibc vs numpy flips "converged at iteration 15" into "never" on the same inputs), so parity with the reference means
parity with **numpy's own float64 routines** on the build the goldens were made with (numpy 2.3.5, x86-64, AVX512_SKX
dispatch).  Neither routine is in `/root/reference`; both live in numpy 2.3.5:

* `np.tanh` → `numpy/_core/src/umath/loops_hyperbolic.dispatch.c.src`, `simd_tanh_f64`: 16 intervals chosen by the exponent
  and first mantissa bit of |x|, a degree-16 polynomial in |x|−b per interval, Horner with fused multiply-adds;
* `np.arctanh` → Intel SVML `__svml_atanh8_ha` (the entry numpy's AVX-512 loop calls; `numpy/_core/src/umath/svml`):
  ½(log(1+|x|) − log(1−|x|)), each logarithm reduced by `VRCP14PD` rounded to 1+4 mantissa bits, a 16-entry hi/lo table of
  log(1+i/16), a degree-9 log1p polynomial and a compensated (two-sum) final accumulation.

`csrc/nlmc_npmath.h` restates both operation for operation (every product-sum is one `fma`, which rounds identically on
x86 FMA3/AVX-512 and in CUDA fp64) with the constants those routines load.  `VRCP14PD` itself is not restated: only its value
rounded to 1+4 mantissa bits is used, and that is a step function of the top 20 mantissa bits of the argument whose 16 steps
were tabulated exhaustively over all 2²⁰ prefixes on a Sapphire Rapids host (the instruction is architecturally
deterministic; the result does not depend on the exponent).  Pinning: the same header compiles as C into the oracle
(`oracle/npmath_host.c`) and is bit-equal to numpy on 9.4e6 arguments per run -- committed sha256 digests per argument set
(`tests/golden/npmath_digests.json`: uniform, log-scale, raw bit patterns, interval edges, reciprocal steps, the `tanh·tanh`
products LBP feeds to `arctanh`), 15k explicit vectors, and live numpy when the host has the golden build's code path
(`tests/test_npmath.py` on CPU, `tests/test_gpu_npmath.py` for the device functions through `nlmc_np_tanh` /
`nlmc_np_arctanh`).

Consequences: K5 reproduces the reference's LBP **bit for bit** (19/19 λ steps of the golden incl. iteration counts and the
divergence point, `test_k5_lbp_golden`, asserted unconditionally); `NMC.run` and `NPT.run` with `doNMC` are bit-exact
free-running (`test_nmc_run_golden`, `test_npt_run_golden_with_nmc_replicas`, and at the native sizes of C1 / C2 against the
oracle's `nmc_run` / `npt_run`); the round-1 backbone override hook and both xfails are gone.  K1 uses the same `tanh` for
non-integer fields, so Gaussian-J decisions are bit-equal too.  On a host whose numpy takes another code path (no AVX-512)
the oracle itself would differ from the goldens; the device would not.

''')

# ------------------------------------------------------------------ data layout: add label mode
s = s.replace('''`u32 thr[n_beta][4]` thresholds (one 16-byte load), `f64 E[n_beta][ladders]`.
''', '''`u32 thr[n_beta][4]` thresholds (one 16-byte load), `f64 E[n_beta][ladders]`.

β-label form of the same layout (`nlmc_msc_create_labelled`; north_star 4): a handle owns the *slots*
[slot_begin, slot_begin+count) of a ladder of `n_beta_total` temperatures -- the whole ladder on one GPU, or one contiguous
block per GPU.  Spins never leave their slot; `u8 labels[n_beta_total][ladders]` (replicated on every rank, 4 KB for C5) says
which temperature each (slot, ladder) is simulated at, `u8 slot_of[...]` is its inverse.  Because the 32 lanes of a word then
sit at different temperatures, the sweep kernel reads its thresholds from **bit planes** rebuilt after every exchange:
`u32 thrbits[6 steps][3 levels][W]` (bit l of word w = bit 31−p of the level-L threshold of lane l; 9 KB for 128 words) and
`u32 thr_lane[W][32][4]` (full thresholds per lane, read only by the ≈1.5 % stragglers).  Random streams are keyed by the
GLOBAL slot and ladder indices, so any partition of the slots evolves bit for bit like the single handle.

Generic engines (K2a/K3): `int8 spins[R][n]` / `bf16 S[R_pad][n_pad]`, one β per row; rows grouped into ladders
(row = ladder·n_beta + slot) with `int32 label[R]`, `slot_of[R]`, `f64 E[R]` for the device-side exchange (`nlmc_exchange.cuh`).
''')

# ------------------------------------------------------------------ K1 exactness: tanh
s = s.replace('''comes from a host LUT computed with numpy's own `tanh`, so ±J decisions are bit-equal; elsewhere CUDA `tanh`
(a 1-ulp difference flips a decision only if the uniform lands in that ulp, ~1e-16 per attempt).''', '''comes from a host LUT computed with numpy's own `tanh`, so ±J decisions are bit-equal; elsewhere the device restatement
of numpy's `tanh` (`nlmc_npmath.h`, table staged in shared memory), bit-equal as well.  A symmetric-J flag computed at
instance creation guards the incremental-field kernel (K1-int pushes J_kj into field j on a flip of k): an asymmetric or
duplicated-entry J -- which the reference accepts through `J.dot(m)` -- takes the general kernel
(`test_k1_asymmetric_j`); the production engines refuse it.''')

# ------------------------------------------------------------------ K3 section numbers
s = s.replace('''sweep 0.44 ms = **9.3e9 attempts/s**.  The update kernel went 193 → 14 µs per block through ncu-driven''', '''sweep 0.355 ms = **1.15e10 attempts/s** (round 1: 0.44 ms).  In round 1 the update kernel went 193 → 14 µs per block through ncu-driven''')
s = s.replace('''shared rows: 0.57 → 0.44 ms per sweep).  NMC phases are per-site modes of the same kernel
(hot backbone at β/temp_x, frozen), and `m_init = M[:, argmin E]` is tracked on the device.''', '''shared rows: 0.57 → 0.44 ms per sweep).  Round 2 (ncu: 6 % of the warp slots, three quarters of the sweep) rewrote it
again: one **warp** per replica; everything off the chain done up front in parallel over the block (one Philox call and four
logits per lane give the thresholds of all 128 sites, NMC phase modes folded in as ±inf / ×temp_x); fully unrolled 8-site
sub-blocks with 128-bit broadcast loads; one CTA per SM (14 replicas) so that the 64 KB coupling block is staged once per
SM: 20.5 → 11.9 µs per block (`profiles/r2_dense_update_kernel_raw.csv`: 3763 → ~2400 warp-instructions per replica-block).
The chain GEMM (10.4 µs per block) + update (11.9 µs) is now evenly split; two ways of overlapping them were built and
measured slower and are kept behind switches: a look-ahead split of K on a second stream (`NLMC_DENSE_LOOKAHEAD`, 0.470 ms:
the third kernel per block and the extra dependency edges cost more than the overlap hides) and programmatic dependent launch
of the chain (`NLMC_DENSE_PDL`, 0.455 ms).  NMC phases are per-site modes of the same kernel (hot backbone at β/temp_x,
frozen), and `m_init = M[:, argmin E]` is tracked on the device.  The heat-bath draw of K2a and K3 uses all 32 random bits with
symmetric tails (a 24-bit uniform rounded to 1.0 used to force a spin down once in 2²⁴ draws; `test_cold_tail_*` at
`global_beta` = 13.6).''')

# ------------------------------------------------------------------ roofline section
s = replace_section(s, "### Why the headline roofline fraction is > 1", "## 5. Measurement (bench.py)", '''### K2 in β-label form, and what the label form costs
`msc_sweep_kernel<kSteps, kPerBit, kSubWarp>`: `kPerBit` takes the threshold bit of every lane from the bit planes (18
extra 16-byte loads per thread, all L1 hits, and three logic operations per word and step where the scalar form has three
FMA-pipe multiply-adds); `kSubWarp` maps several sites to one warp when a site row is shorter than 128 words (a block of a
sharded ladder: 64 / 32 / 16 words at 2 / 4 / 8 GPUs).  K6 in label form is `msc_label_swap_kernel` (one thread per ladder on
the energies of ALL slots, identical on every rank) + `msc_thrbits_kernel` (planes of the local slots) -- 34 µs per round.
ncu (`profiles/r2_sweep_kernel_summary.md`): the scalar form is ALU-pipe bound (67 % active, math-pipe throttle as large as
the load stall); the bit-plane form issues 11 % more instructions at lower occupancy and waits on loads (204 vs 167 µs per
launch).  `bench.py` therefore runs the scalar form where one GPU owns every slot (N = 1) and the bit-plane form where the
temperature range is sharded (N > 1); that 20 % is the whole loss of the strong-scaling curve from N = 1 to N = 2, beyond which
it scales with the slot count (§5).

### The roofline of the headline kernel
SURVEY §8(d) gives two figures: a generic int8 CSR sweep (42 B/attempt at degree 6) and the bit-packed lower bound
(0.25 B/attempt: read the other colour's rows once, write this colour's, 1 bit per spin).  K2 has the second layout, so
`bench.py` reports `roofline` with **0.25 B/attempt**: 134.2 MB per launch ÷ the live launch time (174 µs) = 770 GB/s =
**0.117 of the measured HBM peak** (6545.6 GB/s); ncu measures 110.0 MB of DRAM traffic per launch (0.82 × algorithmic: part
of the 134 MB state stays in the 126 MB L2).  The kernel is not memory bound: `roofline_issue` gives the bound that applies --
117.5 M warp-instructions per launch ÷ launch time = **0.576 of the issue slots** (148 SMs × 4 schedulers × SM clock), with
the ALU pipe at 67 %.  Reaching north_star's "60 % of the memory roofline" would need ≤ 170 warp-instructions per
4096 attempts; six Philox4x32-10 calls alone are 240, so with Philox-10 and an exact 2⁻³² Bernoulli draw the kernel is
instruction bound by construction (round 1 measured every alternative it could think of, `profiles/r1b_sweep_kernel_source.md`;
Philox-7 reaches 3.47e12/s and stays a build option).  The generic 42 B figure is demoted to a note in the JSON line (against
it the same launch would read as ≈20 × the HBM peak, which is what bit-packing buys, not a roofline fraction).

''')

# ------------------------------------------------------------------ measurement
s = replace_section(s, "## 5. Measurement (bench.py)", "## 7. Tests", '''## 5. Measurement (bench.py)

Workload: C5 **as stated** -- 3D ±J EA L=64, 32 β (0.2…2.0) × 128 ladders = 4096 replicas in total; one step = one swap
round (16 sweeps + energies + exchange with 10 pairs per ladder).  Round-2 numbers (`profiles/r2_bench_n{1,2,4,8}.json`, one
8×B200 box, 20 steps, clocks 1965 MHz, no throttle reason; the sustained leg runs the same step for 5.4 s at 700 W):

| GPUs | `value` (strong: 4096 replicas sharded by β) | ms/step | sweep share | efficiency | `weak` (N × 4096 replicas, own ladders) | `e2e` (class API, 20 rounds/call) |
|---|---|---|---|---|---|---|
| 1 | **3.05e12** attempts/s (scalar-threshold kernel, bit exchange) | 5.63 | 0.99 | 1 | -- | 1.3e12 |
| 2 | 4.84e12 (bit-plane kernel, label exchange over NCCL) | 3.55 | 0.95 | 0.79 | 6.08e12 | 1.1e12 |
| 4 | 9.34e12 | 1.84 | 0.94 | 0.77 | 1.215e13 | 1.2e12 |
| 8 | **1.70e13** | 1.01 | 0.93 | 0.70 | **2.43e13** (0.995) | 0.96e12 |

* `value`: CUDA events on the stream the kernels run on, barrier + synchronize on both sides, max over ranks.  The loss from
  1 to 2 GPUs is the label form of the kernel (§4); from 2 to 8 the curve scales with the slot count (×1.93, ×1.82) -- what is
  left is the short-row mapping at 16 words per site and ≈70 µs per round of energies + all-gather + label exchange.
  (Before the energy kernel was rewritten -- §4 K4' -- that fixed part was 1.1 ms per round at 8 GPUs and the step 2.0 ms.)
* `e2e`: `NPT(J, h, mode="production").run(betas, 32, [False]*32, …, num_swap_attempts=K)` with `num_runs = 128`: host scipy J
  in, numpy `M` (float64, last round, run 0: 1.07 GB) and `Energy` out; state resident across the K rounds, per-round
  exchange counts from a device log, the last round recorded on the device in the layout of `M`'s rows and widened
  int8→float64 by a threaded host helper.  One call with K = 20 is 0.25 s: 0.11 s of rounds, 0.05 s instance upload +
  colouring, 0.03 s handle, 0.06 s copy-back + widening -- so `e2e` grows with K (the fixed 0.14 s is per call, not per step)
  and, the fixed part being the same on every rank, does not scale with N.  Under torchrun the same call shards the β range
  and returns the full tuple on every rank.  `e2e_host_buffers` is the round-1 number (one round per call through pinned
  host buffers, the whole packed state both ways: 2.9e12 at N = 1).
* `roofline` / `roofline_issue`: §4.  `traffic` comes from the ncu capture of this build (`profiles/traffic.json`).
* `time_to_target` (BASELINE metric, second half; `tools/time_to_target.py`): parallel tempering to the shipped ground-state
  energy, 16 β, 5 sweeps per round, the reference's pair rule; GPU = 64 ladders at once on the production engine with the
  exchange on the device, CPU = the oracle C port, one ladder on one core; median of 3 seeds:

  | instance | GPU | CPU port | ratio |
  |---|---|---|---|
  | Chimera droplet 128 #001 (real J, fields) | 5.0 ms | 0.83 s | 166 |
  | DCL C8 #00 (463 active spins) | 4.6 ms | 0.57 s | 123 |
  | Wishart planted N=36 α=0.50 #1 | 0.7 ms | 0.67 s | 956 |

* CPU baseline / `--impl reference`: the oracle C port (the reference *algorithm*; the Python reference itself does 4e3
  attempts/s/core, BASELINE.md) at 1.6–2.4e8 attempts/s on 16 host cores.

Other configs (`tools/bench_configs.py`, `profiles/r2_configs_c1_c3.jsonl`, `profiles/r1_configs_c1_c4.jsonl`, one B200):

| config | what | result |
|---|---|---|
| C1 N=800 ±1 graph | `NMC.run` README parameters, sweeps 1e3 (2.48e7 attempts + 10 LBP) | replay (bit-exact, free-running LBP) 8.0 s incl. 2 s of CUDA context; production **0.59 s** (K2a); reference ≈ 50 min |
| C2 EA L=16 | `APT_preprocessor.run` README parameters (37 betas × 100 chains × 1000 sweeps) | 1.45 s |
| C2 EA L=16 | one swap round of 1000 sweeps, 30 betas × 128 ladders, device only | 12.2 ms = **1.29e12 attempts/s** |
| C2 EA L=16 | `NPT.run` with `doNMC` on the 5 coldest (the config as stated), README sweeps, production | **1.07 s** (round 1: 1.5 s): hybrid ladder -- 25 plain replicas bit-packed on K2, 5 NMC replicas on K2a, 50 LBP searches; 0.27 s of it is LBP, 0.3 s the 300 NMC phases, the rest building the 983 MB `M` |
| C2 EA L=16 | `NPT.run` replay (exact), 200 sweeps | 3.2e7 attempts/s (host MT19937 stream bound); with doNMC replicas 1.9e7 |
| C3 SK N=2000 | dense path, 2048 replicas | **1.13e10 attempts/s** (sweep 0.362 ms with the 3-piece J split, 0.341 ms with one piece; round 1: 9.3e9), field GEMM 931 TFLOP/s |
| C4 EA L=32 | `APT_ICM.run` production, 32 β × 10 sub-replicas, 100 rounds × 10 sweeps | 2.0 s (round-1 figure) |
| C5 EA L=64 | the bench line above | |

## 6. Multi-GPU

One process per GPU (`torch.distributed`, NCCL).  Two partitions, both without any movement of spins:

* **β range sharded** (`distributed.ShardedBetaLadder`, the strong-scaling form, north_star 4 / SURVEY 8e): every rank owns a
  contiguous block of the temperature slots of all 128 ladders.  Per round: sweeps of the local slots → bit-sliced energies
  into the all-gather send buffer (`nlmc_msc_energies_dev`) → `all_gather_into_tensor` of one float64 per replica (32 KB in
  total) → `nlmc_msc_exchange_labels`: identical Philox-keyed decisions on every rank (pairs of adjacent *temperatures*, found
  through `slot_of`), labels permuted, the local threshold planes rebuilt.  The handle runs on the caller's stream
  (`nlmc_msc_set_stream`), the one torch orders the collective on, so a round costs no host synchronisation.  Unequal blocks
  (32 = 11+11+10) go through a padded gather.  Proof of equivalence: `tools/multigpu_beta_shard_check.py` under torchrun --
  packed spins, labels and energies of the sharded ensemble equal the single labelled handle **bit for bit** on 2 and on 8
  B200 over NCCL 2.28.9, with 104 / 601 labels sitting across a rank boundary after 8 rounds
  (`profiles/r2_multigpu_beta_shard_check.jsonl`); the same on one GPU for several cuts (`tests/test_gpu_label_exchange.py`)
  and with gloo at world size 2 for the host logic (`tests/test_multigpu_host_gloo.py`).
* **independent ladders** (`distributed.ShardedLadders`, the weak-scaling form of round 1): contiguous blocks of 128 ladders
  per rank, streams keyed by the global ladder index, no exchange step at all; results all-gathered.

`NPT(..., mode="production").run` picks the first form by itself when a process group with more than one rank is initialised.

''')

# ------------------------------------------------------------------ tests
s = replace_section(s, "## 7. Tests", "## 8. Out of scope (SURVEY §2) and why", '''## 7. Tests

`-m "not gpu"` (69 tests, ≈1 min): oracle vs live reference and goldens; the restated `tanh`/`arctanh` against numpy (digests,
vectors, live); host logic of all four classes in replay AND production mode (bit-packed, generic with device exchange,
hybrid ladder) through oracle-backed fakes of the ctypes classes (`tests/fake_backend.py`); header ↔ exported symbols;
no-CPU-fallback; gloo world size 2 for both partitions of the multi-GPU layer.  `-m gpu` (143 tests, ≈1 min on a B200,
**no xfail**): K1/K4/K5/K7 element parity vs goldens and oracle (zeros in the state, scaled/frozen phases, asymmetric J,
n > 200 KiB, `record_from`, empty sweeps) and at the native sizes of C1–C5; whole `NMC.run` on the C1 graph and whole
`NPT.run` with `doNMC` on the C2 lattice against the oracle's run-level restatements; LBP iteration counts and backbones
asserted unconditionally; a randomised sweep over 72 ragged instances; the element-level public methods against the
reference's own; production-path statistics at **3 σ** (exact Boltzmann by full enumeration for K2, K2 in label form, K2a, K3
and the device exchange of K2a/K3; conditional distribution; reference sampler); cold-tail tests at β = 13.6; sharded slots ≡
single handle; hybrid ladder ≡ single-engine ladder; known-answer searches on the planted instances; the bench contract
line; oracle-free properties at the full size of C5.

Memory checking: `compute-sanitizer` is refused on this GPU pool, so out-of-bounds checking was done by review of every
indexed access plus guard tests (sizes that are not multiples of the tile, n = 1, empty rows, rows shorter than a warp).

''')

# ------------------------------------------------------------------ gaps
a = s.index("## 9. Known gaps / next")
s = s[:a] + '''## 9. Known gaps / next
1. K2 in label form costs 20 % per sweep (bit planes instead of scalar thresholds); it is what makes the strong-scaling
   efficiency 0.70–0.79 instead of ≈0.95.  Ideas not yet tried: planes in shared memory for persistent CTAs, or a two-word
   per thread mapping with fewer live registers.
2. K2 (scalar form): ALU-pipe bound at 3.05e12 attempts/s with Philox4x32-10; `profiles/r1b_sweep_kernel_source.md` lists what
   was tried.  Philox-7 (build option) gives 3.47e12.
3. K3: sweep 0.355 ms = 1.15e10 attempts/s (target was 2e10).  The block GEMM (M=2048, N=128, K=2048, 144 CTAs with split-K 9
   and atomic reduction) runs at a third of the stand-alone GEMM's rate and the update kernel is issue bound at 14 warps per
   SM; overlapping the two (look-ahead split, programmatic dependent launch) was measured slower.  The stand-alone GEMM is
   unchanged (931 TFLOP/s, 58 % of burst peak): no persistent scheduler / double-buffered TMEM yet.
4. K2a: one replica per CTA; barrier stall per colour is the limiter (19 colours on C1); native `ATOMS.ADD` restored in round 2
   (verified in SASS) without a measurable change, i.e. the atomics were not the bottleneck.
5. K5: cooperative grid with two grid syncs per iteration (14–20 µs per iteration); LBP calls of different replicas are not
   batched, which is a third of the C2-as-stated run (1.07 s).
6. K1: replay is one warp per chain and latency-bound by construction; a single chain (C1 `NMC.run`) is slower than one CPU
   core running the C port.  It is the parity mode.
7. Production `APT_ICM`: the rounds whose Houdayer edit is observable (the last one, or every round with one sweep per swap)
   move states through the host because K7 takes host buffers; the `find_clusters` growth loop is host-only.
8. `e2e` through the class API carries ≈0.14 s of per-call fixed cost (instance upload, colouring, handle, building the
   float64 `M`), which dominates short calls and is replicated on every rank.
'''
open(p, "w").write(s)
print("ok")
