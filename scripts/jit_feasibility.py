"""Design study, NOT product code and never run on a GPU: what would ONE sweep of a plan look like compiled as
straight-line specialised code instead of being interpreted -- how long does ptxas take on it (a runtime PTX JIT would
pay exactly that), and how many instructions / FP64 instructions / registers does it need?  In-register permutations
become renamings and phases of +-1 / +-i sign flips and swaps.  Folded permutations of the stage boundaries are left
out (address arithmetic only).  Writes /tmp/jit/sweep_<k>.cu; numbers in DESIGN.md section 9.
    python scripts/jit_feasibility.py 0 1 2 5 && nvcc -ptx ... && ptxas -arch=sm_100a -v ..."""
import json, os, sys
OUT = '/tmp/jit'
os.makedirs(OUT, exist_ok=True)
sys.path.insert(0, '.')
import qsyn_b200 as Q
from qsyn_b200 import host, workloads as W

def gen_sweep(sw, n, log2_ncols, name):
    R = 4
    out = []
    A = out.append
    A('#include <cstdint>')
    A('__device__ __forceinline__ void cmulc(double2& a, double wx, double wy){double t=a.x*wx-a.y*wy; a.y=a.x*wy+a.y*wx; a.x=t;}')
    A(f'extern "C" __global__ void __launch_bounds__(128, 5) {name}(double2* __restrict__ u, uint32_t flags) {{')
    A('  extern __shared__ double2 tile[];')
    A('  uint32_t const tid = threadIdx.x, cw = tid & 7u, trow = tid >> 3;')
    tb = sw["tile_bits"]; m = len(tb)
    outer = [b for b in range(n) if b not in tb and not (sw["fixed_one"] >> b) & 1]
    A(f'  uint32_t const tidx = (flags & 1u) ? gridDim.x - 1u - blockIdx.x : blockIdx.x;')
    A(f'  uint32_t const cb = tidx & {(1 << (log2_ncols - 3)) - 1}u, rc = tidx >> {log2_ncols - 3};')
    A(f'  uint32_t row_outer = {sw["fixed_one"]}u;')
    for k, b in enumerate(outer):
        A(f'  row_outer |= ((rc >> {k}) & 1u) << {b};')
    A(f'  size_t const col = ((size_t)cb << 3) + cw;')
    A('  double2 e[16];')
    nst = len(sw["stages"])
    for si, st in enumerate(sw["stages"]):
        rb = st["reg_bits"]
        thr = [b for b in tb if b not in rb]
        # global row of this thread (register bits clear)
        A(f'  {{ // stage {si}')
        A(f'  uint32_t grow = row_outer, lbase = 0;')
        for k, b in enumerate(thr):
            A(f'  grow |= ((trow >> {k}) & 1u) << {b}; lbase |= ((trow >> {k}) & 1u) << {tb.index(b)};')
        if si == 0:
            for i in range(16):
                off = sum(1 << rb[j] for j in range(R) if (i >> j) & 1)
                A(f'  e[{i}] = __ldcg(u + (((size_t)(grow | {off}u)) << {log2_ncols}) + col);')
        else:
            for i in range(16):
                off = sum(1 << tb.index(rb[j]) for j in range(R) if (i >> j) & 1)
                A(f'  e[{i}] = tile[((lbase | {off}u) << 3) + cw];')
            A('  __syncthreads();')
        phys = list(range(16))  # logical -> physical register (in-register permutations are renamings)
        def cond(u):
            c = []
            if u["cother"]: c.append(f'(grow & {u["cother"]}u) == {u["cother"]}u')
            if u["czero"]: c.append(f'(grow & {u["czero"]}u) == 0u')
            return ' && '.join(c)
        for u in st["uops"]:
            k = u["kind"]; j = u["j0"]; c = cond(u)
            pre = f'  if ({c}) {{' if c else '  {'
            if k == "H":
                A(pre)
                for i in range(16):
                    if (i >> j) & 1: continue
                    a, b = phys[i], phys[i | (1 << j)]
                    s = repr(u["c"][0])
                    A(f'    {{ double tx = e[{b}].x * {s}, ty = e[{b}].y * {s}; e[{b}].x = fma({s}, e[{a}].x, -tx); e[{b}].y = fma({s}, e[{a}].y, -ty); e[{a}].x = fma({s}, e[{a}].x, tx); e[{a}].y = fma({s}, e[{a}].y, ty); }}')
                A('  }')
            elif k in ("X", "XC"):
                pairs = []
                for i in range(16):
                    if (i >> j) & 1: continue
                    if k == "XC" and not (i >> u["j1"]) & 1: continue
                    if k == "X" and (i & u["creg"]) != u["creg"]: continue
                    pairs.append((i, i | (1 << j)))
                if not c:
                    for a, b in pairs: phys[a], phys[b] = phys[b], phys[a]   # free
                else:
                    A(pre)
                    for a, b in pairs:
                        A(f'    {{ double2 t = e[{phys[a]}]; e[{phys[a]}] = e[{phys[b]}]; e[{phys[b]}] = t; }}')
                    A('  }')
            elif k == "PHJ":
                wx, wy = u["c"][0], u["c"][1]
                A(pre)
                for i in range(16):
                    if not (i >> j) & 1: continue
                    p = phys[i]
                    if (wx, wy) == (-1.0, 0.0): A(f'    e[{p}].x = -e[{p}].x; e[{p}].y = -e[{p}].y;')
                    elif (wx, wy) == (0.0, 1.0): A(f'    {{ double t = e[{p}].x; e[{p}].x = -e[{p}].y; e[{p}].y = t; }}')
                    elif (wx, wy) == (0.0, -1.0): A(f'    {{ double t = e[{p}].x; e[{p}].x = e[{p}].y; e[{p}].y = -t; }}')
                    else: A(f'    cmulc(e[{p}], {wx!r}, {wy!r});')
                A('  }')
            elif k in ("DTABF", "DTABH"):
                tab = u["table"]
                A(pre)
                idx = 0
                for i in range(16):
                    if k == "DTABH" and not (i >> j) & 1: continue
                    wx, wy = tab[2 * idx], tab[2 * idx + 1]; idx += 1
                    p = phys[i]
                    if (wx, wy) == (1.0, 0.0): continue
                    if (wx, wy) == (-1.0, 0.0): A(f'    e[{p}].x = -e[{p}].x; e[{p}].y = -e[{p}].y;')
                    elif (wx, wy) == (0.0, 1.0): A(f'    {{ double t = e[{p}].x; e[{p}].x = -e[{p}].y; e[{p}].y = t; }}')
                    elif (wx, wy) == (0.0, -1.0): A(f'    {{ double t = e[{p}].x; e[{p}].x = e[{p}].y; e[{p}].y = -t; }}')
                    else: A(f'    cmulc(e[{p}], {wx!r}, {wy!r});')
                A('  }')
            elif k == "SCALN":
                A('  { double sx = 1.0, sy = 0.0;')
                for sc in u["scalars"]:
                    cc = []
                    if sc["cother"]: cc.append(f'(grow & {sc["cother"]}u) == {sc["cother"]}u')
                    if sc["czero"]: cc.append(f'(grow & {sc["czero"]}u) == 0u')
                    d0, d1 = sc["d"][0:2], sc["d"][2:4]
                    sel = sc["sel"]
                    A(f'    if ({" && ".join(cc) if cc else "true"}) {{ double wx = {d0[0]!r}, wy = {d0[1]!r};' + (f' if ((grow >> {sel}) & 1u) {{ wx = {d1[0]!r}; wy = {d1[1]!r}; }}' if sel >= 0 else '') + ' double t = sx * wx - sy * wy; sy = sx * wy + sy * wx; sx = t; }')
                for i in range(16): A(f'    cmulc(e[{i}], sx, sy);')
                A('  }')
            else:
                raise SystemExit("kind " + k)
        last = si == nst - 1
        if last:
            for i in range(16):
                off = sum(1 << rb[j] for j in range(R) if (i >> j) & 1)
                A(f'  __stcg(u + (((size_t)(grow | {off}u)) << {log2_ncols}) + col, e[{phys[i]}]);')
        else:
            for i in range(16):
                off = sum(1 << tb.index(rb[j]) for j in range(R) if (i >> j) & 1)
                A(f'  tile[((lbase | {off}u) << 3) + cw] = e[{phys[i]}];')
            A('  __syncthreads();')
        A('  }')
    A('}')
    return "\n".join(out)

n = 12
qc = host.Circuit(qasm_text=W.random_clifford_t_qasm(n, 2000, 20261019))
d = Q.Plan(n, 2**n, qc.gate_array()).dump()
print(d["sweeps"][0]["stages"][0]["uops"][-1].keys())
which = [int(a) for a in sys.argv[1:]] or [0, 1, 2, 5]
for k in which:
    src = gen_sweep(d["sweeps"][k], n, 12, f"sweep_{k}")
    open(os.path.join(OUT, f"sweep_{k}.cu"), "w").write(src)
    print(k, "stages", len(d["sweeps"][k]["stages"]), "uops", sum(len(s["uops"]) for s in d["sweeps"][k]["stages"]), "lines", src.count("\n"))
