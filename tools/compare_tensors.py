"""Compare two `TENSOR F32 n0 n1 n2 n3` files (localtensor.c:196-239 format) or two PNM images."""
import sys, numpy as np

def load(path):
    with open(path, "rb") as f:
        head = f.readline().split()
        if head[0] == b"TENSOR":
            ne = [int(x) for x in head[2:6]]
            return np.frombuffer(f.read(), dtype=np.float32).reshape(ne[::-1])
        if head[0] in (b"P6", b"P5"):
            toks = head[1:]
            while len(toks) < 3:
                toks += f.readline().split()
            w, h, mx = [int(x) for x in toks[:3]]
            return np.frombuffer(f.read(), dtype=np.uint8).astype(np.float32).reshape(h, w, -1) / 255.0
    raise SystemExit("unknown format: " + path)

a, b = load(sys.argv[1]), load(sys.argv[2])
err = np.abs(a - b).max() / max(np.abs(b).max(), 1e-12)
mse = float(((a - b) ** 2).mean())
psnr = 10 * np.log10(max(np.abs(b).max(), 1e-12) ** 2 / mse) if mse > 0 else float("inf")
print("shape %s  max_rel_err %.4e  psnr %.2f dB  (|b|max %.4g, finite %s)" % (a.shape, err, psnr, np.abs(b).max(), np.isfinite(a).all()))
