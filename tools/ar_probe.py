"""Bring-up probe for the A-resident GEMM variant: decodes which activation rows each tap actually reads."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch, numpy as np
from gpu_util import DEV, lib, ptr, stream
L = lib()
rows, K, N = 300, 32, 32
A = (torch.arange(rows).float().unsqueeze(1) + torch.arange(K).float().unsqueeze(0) / 64.0)
for mode in (0, 1):
    for taps, dil, which in [(2, 8, 0), (2, 8, 1), (2, 1, 1), (2, 3, 1), (3, 5, 2)]:
        W = torch.zeros(taps, N, K)
        W[which] = torch.eye(N)
        pad = dil * (taps - 1) // 2
        out = torch.zeros(rows, N, device=DEV)
        dA, dW, dB = A.to(DEV), W.to(DEV), torch.zeros(N, device=DEV)
        code = L.fs2_op_conv_gemm_ex(stream(), ptr(dA), K, rows, ptr(dW), ptr(dB), taps, dil, K, N,
                                     0, 0.1, None, N, 0, 0, None, None, 0, 0, ptr(out), N)
        torch.cuda.synchronize()
        o = out.cpu()
        want_row = torch.arange(rows).float() + which * dil - pad
        print(f"mode {mode} taps {taps} dil {dil} tap {which} (expect row r{which * dil - pad:+d}): code {code}")
        for r in (0, 1, 7, 8, 9, 40, 127, 128, 200):
            print(f"   r={r:3d} got[0..3]={[round(float(x), 3) for x in o[r, :4]]} got[8..9]={[round(float(x),3) for x in o[r, 8:10]]} want {float(want_row[r]):.0f}")

for cl, DELAY in ((2, 0),):
  L.fs2_debug_set_flag(2, cl)
  print("---- wrong-row map, cluster size", cl, "delay", DELAY)
  for rows in (300, 1000):
      A = (torch.arange(rows).float().unsqueeze(1) + torch.arange(K).float().unsqueeze(0) / 64.0)
      for taps, dil, which in [(2, 8, 0), (2, 8, 1), (2, 1, 1), (2, 2, 1), (2, 3, 1), (2, 4, 1), (2, 5, 1), (2, 6, 1), (2, 7, 1)]:
          W = torch.zeros(taps, N, K); W[which] = torch.eye(N)
          pad = dil * (taps - 1) // 2
          out = torch.zeros(rows, N, device=DEV)
          dA, dW, dB = A.to(DEV), W.to(DEV), torch.zeros(N, device=DEV)
          L.fs2_op_conv_gemm_ex(stream(), ptr(dA), K, rows, ptr(dW), ptr(dB), taps, dil, K, N,
                                0, 0.1, None, N, 0, 0, None, None, 0, 0, ptr(out), N)
          torch.cuda.synchronize()
          o = out.cpu()
          src = torch.arange(rows) + which * dil - pad
          want = torch.where(((src >= 0) & (src < rows)).unsqueeze(1), src.float().unsqueeze(1) + torch.arange(K).float().unsqueeze(0) / 64.0, torch.zeros(1))
          bad = ((o - want).abs() > 0.26).any(1)
          idx = bad.nonzero().flatten().tolist()
          runs, start = [], None
          for i in range(rows + 1):
              b = i < rows and bool(bad[i])
              if b and start is None: start = i
              if not b and start is not None: runs.append((start, i - 1)); start = None
          print(f"rows {rows} taps {taps} dil {dil} tap {which}: {len(idx)} wrong rows, runs {runs[:12]}")
          r0 = runs[0][0] if runs else 0
          print("    sample wrong row", r0, [round(float(x), 2) for x in o[r0, :12]], "want", [round(float(x), 2) for x in want[r0, :12]])
