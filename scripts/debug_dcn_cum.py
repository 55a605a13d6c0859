"""Cumulative error of the neck blocks against the oracle, DCN and plain neck, per precision mode (debugging aid)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from monocon_pytorch_b200 import engine as E      # noqa: E402
from oracle import fixtures as FX                 # noqa: E402
from oracle import monocon_oracle as O            # noqa: E402


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def main():
    dev = torch.device('cuda', 0)
    H, W, B = int(sys.argv[1]), int(sys.argv[2]), 2
    for use_dcn in (True, False):
        sd = FX.make_state_dict(0, use_dcn=use_dcn)
        img = FX.make_images(B, H, W, seed=1)
        pred, inter = O.forward(sd, img, return_intermediates=True)
        for precision in ('fp32_simt', 'fp32'):
            eng = E.Engine(dev, B, H, W, precision, use_dcn=use_dcn)
            eng.load_state_dict(sd)
            if eng.tensor_core_fp32:
                eng.calibrate_scales(img.to(dev))
            out = eng.forward(img.to(dev))
            torch.cuda.synchronize()
            line = [f'l{l} {rel(eng.debug_tensor(f"backbone.level{l}", B).cpu(), inter["backbone"][l]):.1e}' for l in range(2, 6)]
            for name, t in inter['neck'].items():
                line.append(f'{name[5:]} {rel(eng.debug_tensor(name, B).cpu(), t):.1e}')
            line.append('maps ' + ' '.join(f'{rel(t.cpu(), pred[k]):.1e}' for k, t in zip(E.PRED_NAMES, out)))
            print(f'dcn={use_dcn} {precision}: ' + ' | '.join(line), flush=True)
            eng.close()


if __name__ == '__main__':
    main()
