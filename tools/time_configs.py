import torch, time, sys
sys.path.insert(0,'/root/repo')
import speech_enhancement_by_s3prl_b200 as se
from speech_enhancement_by_s3prl_b200 import synth
dev=torch.device('cuda',0)
for (nfreq,win,hop) in [(257,32,16),(201,25,10),(513,64,16)]:
    pre=se.OnlinePreprocessor(sample_rate=16000,win_ms=win,hop_ms=hop,n_freq=nfreq).to(dev); pre.channel_inp,pre.channel_tar=0,1
    torch.manual_seed(1337); head=se.LinearResidual(input_size=nfreq,output_size=nfreq).to(dev)
    eng=se.EnhancementEngine(pre,head,log_features=True,precision=1)
    lengths,wavs=synth.batch(64,4.0); lengths,wavs=lengths.to(dev),wavs.to(dev)
    g=eng.capture_bound(lengths,wavs)
    for _ in range(5): g["graph"].replay()
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): g["graph"].replay()
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/50
    print(nfreq,win,hop,'ms/step',ms,'audio-s/s',256/ms*1e3, 'sisdr', g["sisdr"].mean().item())
