#!/usr/bin/env python3
"""Two-pass split N = R1 * R2 for every even image edge with prime factors 2 / 3 / 5 / 7 that has no hand-tuned entry in
fft_regs.cuh: prints the BFFT_GEO_AUTO lines (pasted into fft_regs.cuh) and the X(n) list for BIOEM_SIZES_AUTO.

Rules (the fused kernel's lane roles): R1 even (the packed layout pairs sub-sequences), R2 <= 32 (one lane per
sub-sequence in the first radix pass), KC = largest divisor of N/2 with KC * R2 <= 32 columns per warp task; prefer full
warps in both passes and radices <= 24 (registers)."""
import sys

TUNED = {32, 36, 48, 64, 96, 128, 160, 192, 224, 256, 288, 320, 360, 384, 400,
         100, 120, 144, 200, 216, 240, 300, 336, 420, 432, 448, 480, 500, 512}


def smooth(n):
    for p in (2, 3, 5, 7):
        while n % p == 0:
            n //= p
    return n == 1


def pick(n):
    best = None
    for r1 in range(2, 33, 2):
        if n % r1:
            continue
        r2 = n // r1
        if r2 > 32 or r2 < 2:
            continue
        kc = max(d for d in range(1, max(1, 32 // r2) + 1) if (n // 2) % d == 0)
        p2 = -(-(kc * r1) // 32)
        if p2 * r2 > 32:          # validity mask of the generic variant: P2 * NK bits, NK = R2
            continue
        eff1, eff2 = kc * r2 / 32, kc * r1 / (32 * p2)
        score = 0.65 * eff1 + 0.35 * eff2 - 0.03 * max(0, r1 - 20) - 0.03 * max(0, r2 - 25)
        if best is None or score > best[0]:
            best = (score, r1, r2)
    return best


lines, names = [], []
for n in range(16, 513, 2):
    if not smooth(n) or n in TUNED:
        continue
    b = pick(n)
    if b is None:
        print(f"// {n}: no valid split", file=sys.stderr)
        continue
    _, r1, r2 = b
    fkc = max(d for d in range(1, 25) if (n // 2) % d == 0)
    pc = max(1, min(16, 256 // max(r1, r2)))
    lines.append(f"BFFT_GEO_AUTO({n}, {r1}, {r2}, {fkc}, {pc})")
    names.append(f"X({n})")
print("\n".join(lines))
print("#define BIOEM_SIZES_AUTO(X) " + " ".join(names))
