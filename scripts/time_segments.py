"""Where does an iteration go?  Captures the three parts of TrainStep (critic step, Adam(D) + generator step,
Adam(G) + statistics) as separate graphs on one GPU and times each with CUDA events.
usage (B200): python scripts/time_segments.py [res] [alpha] [batch]"""
import sys
import torch
sys.path.insert(0, '.')
from neuron_gan_b200.train_step import TrainStep, build_networks
from oracle import pggan_oracle as O     # synthetic images only

res, alpha, B = int(sys.argv[1]) if len(sys.argv) > 1 else 512, float(sys.argv[2]) if len(sys.argv) > 2 else 1.0, \
    int(sys.argv[3]) if len(sys.argv) > 3 else 16
G, D = build_networks(res, alpha, seed=1, device='cuda')
step = TrainStep(G, D)
step.segment_graphs = True
xs = [O.synthetic_images(B, res, seed=100 + i).cuda() for i in range(4)]
draws = [step.draw(B, 'cuda') for _ in range(4)]
for i in range(4):
    step(xs[i % 4], draws[i % 4])
step.segment_events = []
for i in range(20):
    step(xs[i % 4], draws[i % 4])
torch.cuda.synchronize()
t = torch.tensor([[e[k].elapsed_time(e[k + 1]) for k in range(3)] for e in step.segment_events]).mean(0)
print(f'res={res} alpha={alpha} B={B}: critic step {t[0]:.3f} ms, Adam(D)+generator step {t[1]:.3f} ms, '
      f'Adam(G)+stats {t[2]:.3f} ms, total {t.sum():.3f} ms')
