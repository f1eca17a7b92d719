import torch, sys
sys.path.insert(0,'/root/repo')
import speech_enhancement_by_s3prl_b200 as se
from speech_enhancement_by_s3prl_b200 import synth
dev=torch.device('cuda',0)
nfreq,win,hop=int(sys.argv[1]),int(sys.argv[2]),int(sys.argv[3])
pre=se.OnlinePreprocessor(sample_rate=16000,win_ms=win,hop_ms=hop,n_freq=nfreq).to(dev); pre.channel_inp,pre.channel_tar=0,1
torch.manual_seed(1337); head=se.LinearResidual(input_size=nfreq,output_size=nfreq).to(dev)
eng=se.EnhancementEngine(pre,head,log_features=True,precision=1)
lengths,wavs=synth.batch(64,4.0); lengths,wavs=lengths.to(dev),wavs.to(dev)
for _ in range(3): out=eng.eval_step(lengths,wavs)
torch.cuda.synchronize()
print(out["sisdr"].mean().item())
