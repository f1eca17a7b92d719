import sys; sys.path.insert(0,"/root/repo")
import torch, speech_enhancement_by_s3prl_b200 as se
from speech_enhancement_by_s3prl_b200 import synth
dev=torch.device("cuda",0)
pre=se.OnlinePreprocessor(sample_rate=16000,win_ms=32,hop_ms=16,n_freq=257).to(dev); pre.channel_inp,pre.channel_tar=0,1
torch.manual_seed(1337); head=se.LinearResidual(input_size=257,output_size=257,precision=1).to(dev)
eng=se.EnhancementEngine(pre,head,precision=1); opt=torch.optim.Adam(head.parameters(),lr=1e-4)
lengths,wavs=synth.batch(64,4.0); lengths,wavs=lengths.to(dev),wavs.to(dev)
for _ in range(3): eng.train_step(lengths,wavs,se.SISDR(),optimizer=opt,grad_clip=1.0)
torch.cuda.synchronize()
