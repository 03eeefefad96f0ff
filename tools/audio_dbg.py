import numpy as np, torch, sys
sys.path.insert(0,'.')
from oracle import driver_oracle
from tdc_video_b200.audio import pool_audio_per_frame
n_frames=27; rs=np.random.RandomState(3); seconds,flags=n_frames,[1]*n_frames
n_win=(seconds+9)//10; windows=[]
for w in range(n_win):
    tlen=min(10,seconds-10*w)*50-(7 if w==n_win-1 else 0)
    windows.append(torch.from_numpy(rs.standard_normal((1,tlen,768)).astype(np.float32)))
ref=driver_oracle.audio_frames_from_beats(windows,flags,n_frames)
got=pool_audio_per_frame([w.cuda().bfloat16() for w in windows],flags,n_frames).float().cpu()
d=(got-ref).abs()
print("max diff",d.max().item(),"argmax",np.unravel_index(d.argmax().item(),d.shape))
print("per-frame max diff",[round(x,3) for x in d.amax(dim=(1,2)).tolist()])
print("zeros got",int((got==0).sum()),"zeros ref",int((ref==0).sum()))
