"""CUDA source generation for the persistent fused Scan kernel (K7) — elementwise recurrences.

One launch runs ALL `n_steps` iterations: each thread owns one element of the carried state(s), keeps the tap windows
in registers for the whole time loop (the inner graph is inlined as a ScalarProgram), streams sequence slices in with
coalesced loads (next step prefetched), and only writes the steps that survive in the (possibly truncated, circular)
trace buffers.  Replaces T x (Cython cell shuffling + inner CVM call) of pytensor/scan/scan_perform.pyx:311-541.

Buffer protocol (identical to the general path, see vm/nodes_scan.py): state buffer `b` has `store` slots along dim 0,
the first L = -mintap slots hold the initial taps; step i writes slot (L + i) mod store; a value survives iff
i >= T - store.  Rotation into chronological order happens after the kernel (host-issued copies).
"""

from __future__ import annotations

from .elemwise import MAX_DIMS
from .scalar import CTYPE, PRELUDE, ScalarProgram, emit_body


def gen_fused_scan_kernel(prog: ScalarProgram, name: str, n_seq: int, state_taps, n_nit: int, n_nonseq: int) -> str:
    """prog inputs: [seq_0.., (state_0 taps in the op's tap order).., nonseq_0..]; outputs: [state_0 new.., nit_0..].

    state_taps: list of tuples of (negative) taps per recurrent state.
    Operand order of the kernel (pointers and ScDims.st rows): seqs, state buffers, nit buffers, non-seqs."""
    n_state = len(state_taps)
    nops = n_seq + n_state + n_nit + n_nonseq
    seq_dt = prog.in_dtypes[:n_seq]
    pos = n_seq
    state_dt = []
    for taps in state_taps:
        state_dt.append(prog.in_dtypes[pos])
        pos += len(taps)
    nonseq_dt = prog.in_dtypes[pos:pos + n_nonseq]
    nit_dt = prog.out_dtypes[n_state:n_state + n_nit]
    params = []
    for k, d in enumerate(seq_dt):
        params.append(f"const {CTYPE[d]}* __restrict__ pseq{k}")
    for k, d in enumerate(state_dt):
        params.append(f"{CTYPE[d]}* pst{k}")
    for k, d in enumerate(nit_dt):
        params.append(f"{CTYPE[d]}* __restrict__ pnit{k}")
    for k, d in enumerate(nonseq_dt):
        params.append(f"const {CTYPE[d]}* __restrict__ pns{k}")
    params += ["const ScDims d", "long long total", "long long T"]

    L = [-min(t) for t in state_taps]
    lines = []
    A = lines.append
    A(PRELUDE)
    A(emit_body(prog))
    A(f"struct ScDims {{ int ndim; long long shape[{MAX_DIMS}]; long long st[{max(nops, 1)}][{MAX_DIMS}]; "
      f"long long tstride[{max(nops, 1)}]; long long store[{max(n_state + n_nit, 1)}]; }};")
    A(f'extern "C" __global__ void __launch_bounds__(256) {name}({", ".join(params)}) {{')
    A("  const long long gstride = (long long)gridDim.x * blockDim.x;")
    A("  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gstride) {")
    A("    long long rem = e;")
    for j in range(nops):
        A(f"    long long off{j} = 0;")
    A("#pragma unroll")
    A(f"    for (int k = {MAX_DIMS} - 1; k >= 0; --k) {{")
    A("      if (k < d.ndim) {")
    A("        const long long q = rem / d.shape[k]; const long long c = rem - q * d.shape[k]; rem = q;")
    for j in range(nops):
        A(f"        off{j} += c * d.st[{j}][k];")
    A("      }")
    A("    }")
    o_seq, o_st, o_nit, o_ns = 0, n_seq, n_seq + n_state, n_seq + n_state + n_nit
    for k, dt in enumerate(nonseq_dt):
        A(f"    const {CTYPE[dt]} ns{k} = pns{k}[off{o_ns + k}];")
    for k, dt in enumerate(state_dt):
        for j in range(L[k]):
            A(f"    {CTYPE[dt]} w{k}_{j} = pst{k}[off{o_st + k} + {j}LL * d.tstride[{o_st + k}]];")
    for k, dt in enumerate(seq_dt):
        A(f"    {CTYPE[dt]} sq{k} = (T > 0) ? pseq{k}[off{o_seq + k}] : ({CTYPE[dt]})0;")
    # The time loop runs in two phases.  Phase 1 covers the steps whose results do not survive in any (truncated) trace
    # buffer: body only, no stores, no slot bookkeeping — for `hs[-1]` graphs (store = 2) that is all but the last steps.
    # Phase 2 stores: slot = (L + i) % store is computed ONCE at its start and then advanced incrementally together with
    # the slot's offset (no 64-bit modulo / multiply per step; a full-trace Scan spends all its steps here).
    n_out = n_state + n_nit
    A("    const int Ti = (int)T;")
    A("    int i0 = Ti;")
    for k in range(n_state):
        A(f"    const int fs{k} = (int)max(0LL, T - d.store[{k}]); i0 = min(i0, fs{k});")
    for k in range(n_nit):
        A(f"    const int fn{k} = (int)max(0LL, T - d.store[{n_state + k}]); i0 = min(i0, fn{k});")

    def step(stores: bool):
        for k, dt in enumerate(seq_dt):
            A(f"      const {CTYPE[dt]} cur_sq{k} = sq{k};")
            A(f"      if (i + 1 < Ti) sq{k} = pseq{k}[off{o_seq + k} + (long long)(i + 1) * d.tstride[{o_seq + k}]];")
        args = [f"cur_sq{k}" for k in range(n_seq)]
        for k, taps in enumerate(state_taps):
            for t in taps:
                args.append(f"w{k}_{L[k] + t}")
        args += [f"ns{k}" for k in range(n_nonseq)]
        for k, dt in enumerate(state_dt):
            A(f"      {CTYPE[dt]} nv{k};")
        for k, dt in enumerate(nit_dt):
            A(f"      {CTYPE[dt]} nn{k};")
        outs = [f"nv{k}" for k in range(n_state)] + [f"nn{k}" for k in range(n_nit)]
        A(f"      ptk_body({', '.join(args + outs)});")
        for k in range(n_state):
            for j in range(L[k] - 1):
                A(f"      w{k}_{j} = w{k}_{j + 1};")
            A(f"      w{k}_{L[k] - 1} = nv{k};")
        if not stores:
            return
        for k in range(n_state):
            cond = f"if (i >= fs{k}) " if n_out > 1 else ""
            A(f"      {cond}pst{k}[off{o_st + k} + wo{k}] = nv{k};")
            A(f"      wo{k} += d.tstride[{o_st + k}];")
            A(f"      if (++sl{k} == st{k}) {{ sl{k} = 0; wo{k} = 0; }}")
        for k in range(n_nit):
            cond = f"if (i >= fn{k}) " if n_out > 1 else ""
            A(f"      {cond}pnit{k}[off{o_nit + k} + no{k}] = nn{k};")
            A(f"      no{k} += d.tstride[{o_nit + k}];")
            A(f"      if (++sn{k} == stn{k}) {{ sn{k} = 0; no{k} = 0; }}")

    A("    for (int i = 0; i < i0; ++i) {")
    step(False)
    A("    }")
    if n_out == 1 and n_state == 1:
        # single recurrent output (the common case, incl. BASELINE configs[3]): walk the trace buffer in RUNS that end
        # where the circular slot index wraps — inside a run a step costs body + one store + one pointer bump
        A("    const int st0 = (int)d.store[0];")
        A(f"    int sl0 = (int)(({L[0]}LL + i0) % d.store[0]);")
        A(f"    {CTYPE[state_dt[0]]}* wp = pst0 + off{o_st} + sl0 * d.tstride[{o_st}];")
        A("    for (int i = i0; i < Ti;) {")
        A("      const int run = min(Ti - i, st0 - sl0);")
        A("      const int i_end = i + run;")
        A("      for (; i < i_end; ++i) {")
        step(False)
        A("        *wp = nv0;")
        A(f"        wp += d.tstride[{o_st}];")
        A("      }")
        A("      sl0 += run;")
        A(f"      if (sl0 == st0) {{ sl0 = 0; wp = pst0 + off{o_st}; }}")
        A("    }")
    else:
        for k in range(n_state):
            A(f"    const int st{k} = (int)d.store[{k}];")
            A(f"    int sl{k} = (int)(({L[k]}LL + i0) % d.store[{k}]);")
            A(f"    long long wo{k} = sl{k} * d.tstride[{o_st + k}];")
        for k in range(n_nit):
            A(f"    const int stn{k} = (int)d.store[{n_state + k}];")
            A(f"    int sn{k} = stn{k} > 0 ? (int)(i0 % d.store[{n_state + k}]) : 0;")
            A(f"    long long no{k} = sn{k} * d.tstride[{o_nit + k}];")
        A("    for (int i = i0; i < Ti; ++i) {")
        step(True)
        A("    }")
    A("  }")
    A("}")
    return "\n".join(lines)
