#!/usr/bin/env python
"""Shared-memory bank-conflict simulator for the FFT exchange patterns (16-byte elements, 128-bit
accesses are served per quarter-warp: 8 lanes, conflict-free iff their 16-byte slots differ mod 8)."""
import itertools
import sys


def plan(N):
    M = N // 16
    R2 = 16 if M >= 16 else M
    R3 = N // (16 * R2)
    return M, R2, R3


def accesses(N, T):
    """(name, fn(slot b) -> element addresses, one per exchange instruction) for the kernel's passes.

    3-pass plans: pass 2 works in place on each thread's own 16 locations and pass 3 only reads what
    the R3 neighbouring slots (same warp) wrote — see fft_pencil in csrc/zplt_fft.cuh."""
    M, R2, R3 = plan(N)
    out = [("p1 write", lambda b, M=M: [k * M + b for k in range(16)])]
    if R3 == 1:
        L, NB = 1, 16 // R2

        def rd(b, R=R2, NB=NB, M=M):
            res = []
            for j in range(NB):
                q = b + j * M
                for n in range(R):
                    res.append(q * R + n)
            return res
        out.append(("p2 read", rd))
        return out

    def p2(b, R3=R3):
        k1, i = divmod(b, R3)
        return [k1 * 16 * R3 + n * R3 + i for n in range(16)]
    out.append(("p2 read/write", p2))

    def p3(b, R3=R3):
        k1, i = divmod(b, R3)
        res = []
        for j in range(16 // R3):
            for n in range(R3):
                res.append(k1 * 16 * R3 + (i + R3 * j) * R3 + n)
        return res
    out.append(("p3 read", p3))
    return out


def conflicts(N, T, pstride, swz=lambda a: a):
    M, _, _ = plan(N)
    NT = T * M
    total, ideal = 0, 0
    detail = []
    for name, fn in accesses(N, T):
        lists = {tid: fn(tid // T) for tid in range(NT)}
        ninstr = len(lists[0])
        w = 0
        for q0 in range(0, NT, 8):
            lanes = list(range(q0, min(q0 + 8, NT)))
            for i in range(ninstr):
                banks = {}
                for tid in lanes:
                    p = tid % T
                    slot = p * pstride + swz(lists[tid][i])
                    banks.setdefault(slot % 8, set()).add(slot)
                w += max(len(v) for v in banks.values())
        nq = (NT + 7) // 8
        detail.append((name, w / (nq * ninstr)))
        total += w
        ideal += nq * ninstr
    return total / ideal, detail


def conflicts64(N, T, pstride, swz=lambda a: a):
    """The same exchange patterns with 8-byte elements (the exchange split into a real and an imaginary round — DESIGN.md §9.1):
    a 64-bit access is served per half-warp, 16 lanes, conflict-free iff their 8-byte words differ mod 16."""
    M, _, _ = plan(N)
    NT = T * M
    total, ideal, detail = 0, 0, []
    for name, fn in accesses(N, T):
        lists = {tid: fn(tid // T) for tid in range(NT)}
        ninstr = len(lists[0])
        w = 0
        for q0 in range(0, NT, 16):
            lanes = list(range(q0, min(q0 + 16, NT)))
            for i in range(ninstr):
                banks = {}
                for tid in lanes:
                    word = (tid % T) * pstride + swz(lists[tid][i])
                    banks.setdefault(word % 16, set()).add(word)
                w += max(len(v) for v in banks.values())
        nq = (NT + 15) // 16
        detail.append((name, w / (nq * ninstr)))
        total += w
        ideal += nq * ninstr
    return total / ideal, detail


def search64(N, T):
    """Best (ratio, pencil-stride offset, xor shift) for the 8-byte exchange; shift 0 = no swizzle."""
    best = None
    for ps_off in range(0, 17):
        for sh in range(0, 9):
            for mask in (1, 3, 7, 15):
                swz = (lambda a: a) if sh == 0 else (lambda a, sh=sh, mask=mask: a ^ ((a >> sh) & mask))
                r, d = conflicts64(N, T, N + ps_off, swz)
                if best is None or r < best[0] - 1e-9:
                    best = (r, ps_off, sh, mask if sh else 0, d)
                if sh == 0:
                    break
    return best


if __name__ == "__main__":
    if len(sys.argv) > 3 and sys.argv[3] == "split":
        r, off, sh, mask, d = search64(int(sys.argv[1]), int(sys.argv[2]))
        print(f"8-byte exchange N={sys.argv[1]} T={sys.argv[2]}: best PSTRIDE=N+{off}, swizzle a^((a>>{sh})&{mask}) : "
              f"avg wavefronts/ideal {r:.3f}  " + " ".join(f"{n}:{v:.2f}" for n, v in d))
        sys.exit(0)
    N = int(sys.argv[1])
    T = int(sys.argv[2])
    for ps_off in range(0, 9):
        r, d = conflicts(N, T, N + ps_off)
        print(f"N={N} T={T} PSTRIDE=N+{ps_off}: avg wavefronts/ideal {r:.2f}  " + " ".join(f"{n}:{v:.2f}" for n, v in d))
    # xor swizzle candidates
    for sh in (2, 3, 4, 5, 6):
        for ps_off in (0, 1, 2, 4):
            r, d = conflicts(N, T, N + ps_off, lambda a, sh=sh: a ^ ((a >> sh) & 7))
            print(f"N={N} T={T} PSTRIDE=N+{ps_off} swz a^((a>>{sh})&7): {r:.2f}  " + " ".join(f"{n}:{v:.2f}" for n, v in d))


def search(N, T):
    best = None
    for ps_off in range(0, 9):
        for sh in (0, 1, 2, 3, 4, 5, 6, 7):
            swz = (lambda a: a) if sh == 0 else (lambda a, sh=sh: a ^ ((a >> sh) & 7))
            r, d = conflicts(N, T, N + ps_off, swz)
            if best is None or r < best[0] - 1e-9:
                best = (r, ps_off, sh)
    return best
