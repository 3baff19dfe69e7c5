"""Offline fit of tools/texvalue_probe.cu's dumps: how far is the texture unit's bilinear result from the oracle's model
(weights snapped to 1/256, fp32 lerps of texel / 255 or of the sRGB formula), and which refinement explains the rest?
usage: python tools/texvalue_fit.py [dir]   (dir holds texvalue_unorm.bin, texvalue_srgb.bin, texvalue_srgb_table.bin)"""
import sys
import numpy as np

d = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out"
W = 64


def load(name):
    raw = open(f"{d}/{name}", "rb").read()
    tex = np.frombuffer(raw[:W * W * 4], np.uint8).reshape(W, W, 4)
    rec = np.frombuffer(raw[W * W * 4:], np.float32).reshape(-1, 6)
    return tex, rec[:, :2].astype(np.float32), rec[:, 2:]


def srgb_formula(i):
    c = np.arange(256, dtype=np.float32) / np.float32(255)
    return np.where(c <= 0.04045, c / np.float32(12.92), np.power((c + np.float32(0.055)) / np.float32(1.055), np.float32(2.4))).astype(np.float32)[i]


def model(tex, uv, decode, snap_bits=8, order="x_first", quant=None):
    x = uv[:, 0] * np.float32(W) - np.float32(0.5)
    y = uv[:, 1] * np.float32(W) - np.float32(0.5)
    s = np.float32(1 << snap_bits)
    x = np.floor(x * s + np.float32(0.5)) / s
    y = np.floor(y * s + np.float32(0.5)) / s
    fx, fy = np.floor(x), np.floor(y)
    ax, ay = (x - fx).astype(np.float32)[:, None], (y - fy).astype(np.float32)[:, None]
    x0, y0 = fx.astype(int) % W, fy.astype(int) % W
    x1, y1 = (x0 + 1) % W, (y0 + 1) % W
    t00, t10, t01, t11 = (decode(tex[yy, xx]) for yy, xx in ((y0, x0), (y0, x1), (y1, x0), (y1, x1)))
    if order == "x_first":
        top = t00 + ax * (t10 - t00)
        bot = t01 + ax * (t11 - t01)
        if quant:
            top, bot = np.round(top * quant) / quant, np.round(bot * quant) / quant
        out = top + ay * (bot - top)
    else:  # weighted sum of four
        out = t00 * (1 - ax) * (1 - ay) + t10 * ax * (1 - ay) + t01 * (1 - ax) * ay + t11 * ax * ay
    return out.astype(np.float32)


def report(label, got, want):
    e = np.abs(got.astype(np.float64) - want.astype(np.float64))
    n4 = len(e) // 4
    print(f"   {label:46s} on-grid max {e[:n4].max():.3e} q99 {np.quantile(e[:n4], .99):.3e} | anywhere max {e[n4:].max():.3e} q99 {np.quantile(e[n4:], .99):.3e} median {np.median(e[n4:]):.3e}")


tex, uv, got = load("texvalue_unorm.bin")
print("unorm:")
unorm = lambda t: t.astype(np.float32) / np.float32(255)
report("8-bit weights, fp32 lerp x then y", got, model(tex, uv, unorm))
report("8-bit weights, weighted sum of four", got, model(tex, uv, unorm, order="sum"))
for bits in (7, 9, 10, 12):
    report(f"{bits}-bit weights", got, model(tex, uv, unorm, snap_bits=bits))
# is the output itself quantised?
for q in (255 * 256, 255 * 65536, 1 << 16, 1 << 24):
    r = got[:, :3].astype(np.float64) * q
    print(f"   output * {q}: max distance from an integer {np.abs(r - np.round(r)).max():.4f}")

tex, uv, got = load("texvalue_srgb.bin")
tab = np.frombuffer(open(f"{d}/texvalue_srgb_table.bin", "rb").read(), np.float32).reshape(256, 2)
print("srgb table: max |hardware - formula| = %.3e (alpha channel linear: max |a - i/255| = %.3e)" %
      (np.abs(tab[:, 0] - srgb_formula(np.arange(256))).max(), np.abs(tab[:, 1] - np.arange(256, dtype=np.float32) / 255).max()))
i = np.argmax(np.abs(tab[:, 0] - srgb_formula(np.arange(256))))
print("   worst code", i, "hardware", tab[i, 0], "formula", srgb_formula(np.arange(256))[i])


def srgb_dec(table):
    def f(t):
        out = table[t].astype(np.float32)
        out[..., 3] = t[..., 3].astype(np.float32) / np.float32(255)
        return out
    return f


print("srgb:")
report("formula, 8-bit weights", got[:, :3], model(tex, uv, srgb_dec(srgb_formula(np.arange(256))))[:, :3])
report("hardware table, 8-bit weights", got[:, :3], model(tex, uv, srgb_dec(tab[:, 0]))[:, :3])
