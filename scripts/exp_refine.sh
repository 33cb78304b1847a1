timeout 300 python - <<'PY'
import torch, time, sys
sys.path.insert(0,'.')
import vision_spectra_b200 as pkg
from vision_spectra_b200.sweep import CheckpointLayout, SweepRunner
dev=torch.device('cuda',0); lay=CheckpointLayout.vit(192,6); eng=pkg.SpectraEngine(dev); r=SweepRunner(eng,lay)
g=torch.Generator(device=dev).manual_seed(1)
arenas=[torch.randn(lay.arena_elems,generator=g,device=dev)*0.02 for _ in range(93)]
def run(tag):
    for _ in range(3): r.run_device(arenas)
    torch.cuda.synchronize()
    sm=[]; res=r.run_device(arenas, stage_ms=sm); rec=res.records_host()
    print(tag, [round(x,2) for x in sm], 'refined', int((rec['status']&32).astype(bool).sum()))
run('legacy default stream')
s=torch.cuda.Stream()
with torch.cuda.stream(s):
    run('user stream')
PY
