import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, pkb200
pk = pkb200.pk
torch.cuda.set_device(0)
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
for lut in (True, False):
    code = pk.Code(6, 6, device=0)
    code.set_lut(lut)
    kan = pk.Kaneko(code, J=15)
    B = 16384
    y = torch.empty((B, 63), dtype=torch.float64, device="cuda")
    kan.generate_frames_dev(5.0, 10, 1, 0, B, y.data_ptr(), stream=st.cuda_stream)
    torch.cuda.synchronize()
    yy = y[15201:15202].contiguous()
    yh = yy.cpu().numpy()
    dec, tr, recs, tot = kan.decode(yh)
    print("lut", lut, "trials", tr, "flags", recs["flags"], "extra", recs["extra_cmp"], recs["extra_sum"])
    d = torch.zeros((1, 63), dtype=torch.uint8, device="cuda"); t = torch.zeros(1, dtype=torch.int32, device="cuda"); to = torch.zeros(8, dtype=torch.int64, device="cuda")
    for r in range(3):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(st); kan.decode_dev(yy.data_ptr(), 1, d.data_ptr(), t.data_ptr(), None, to.data_ptr(), st.cuda_stream); e1.record(st); torch.cuda.synchronize()
        print("   ms", e0.elapsed_time(e1))
    for lim in (32, 1024, 4096):
        kan.set_phase_a_limit(lim)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(st); kan.decode_dev(yy.data_ptr(), 1, d.data_ptr(), t.data_ptr(), None, to.data_ptr(), st.cuda_stream); e1.record(st); torch.cuda.synchronize()
        print("   limit_a", lim, "ms", e0.elapsed_time(e1))
    np.save("gpurun_out/frame15201.npy", yh)
