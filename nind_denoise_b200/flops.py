"""Algorithmic work of the networks (2*MAC; Conv2d counts output pixels, ConvTranspose2d input pixels —
SURVEY §8d / BASELINE.md §3).  Used by bench.py for the roofline figures."""


def utnet_flops(cs: int, funit: int = 64) -> float:
    f = funit
    sizes = []  # (MACs)
    e = cs
    macs = 27 * f * (cs + 2) ** 2 + 9 * f * f * cs ** 2            # convs1.0, convs1.2
    c = f
    pooled = []
    for _ in range(3):                                             # convs2..4
        p = e // 2
        pooled.append(p)
        macs += 9 * c * 2 * c * (p - 2) ** 2 + 9 * 2 * c * 2 * c * (p - 4) ** 2
        e, c = p - 4, 2 * c
    p4 = e // 2
    macs += 9 * c * 2 * c * (p4 - 2) ** 2                          # bottom.0  (c = 8f)
    macs += 9 * 2 * c * 2 * c * (p4 - 2) ** 2                      # bottom.2  (ConvT: input pixels)
    width, s = 2 * c, p4                                           # decoder
    for _ in range(4):
        half = width // 2
        macs += 4 * width * half * s ** 2                          # up (2x2, input pixels)
        s2 = 2 * s
        macs += 9 * width * half * s2 ** 2                         # tconvs.0 on the concat
        macs += 9 * half * half * (s2 + 2) ** 2                    # tconvs.2
        width, s = half, s2 + 4
    macs += 3 * f * (cs + 4) ** 2                                  # 1x1 head
    return 2.0 * macs


def unet_flops(cs: int) -> float:
    return 256.02e9 * (cs / 512.0) ** 2
