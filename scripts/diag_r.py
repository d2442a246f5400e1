import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, 'tests')
import numpy as np, torch
from geneface_b200 import synthetic, utils
from oracle import field as OF
Himg = 32
model, hp = synthetic.build_model(torso=False, bitfield='R', seed=3, sigma_scale=4.0)
sd = synthetic.state_to_numpy(model)
fi = synthetic.frame_inputs(Himg, Himg)
with torch.no_grad():
    cf = model.cal_cond_feat(fi['cond'])
out = model.render_fused(cf, Himg, Himg, pose=fi['pose'][0], intrinsics=fi['intrinsics'], bg_color=fi['bg_color'], dt_gamma=hp['dt_gamma'],
                         max_steps=hp['max_steps'], precision='fp32', want=('weights_sum', 'n_samples', 'counters', 'term_hist'))
torch.cuda.synchronize()
ro, rd = OF.get_rays(fi['pose'][0].cpu().numpy(), fi['intrinsics'], Himg, Himg)
fo = OF.FieldOracle(sd, bound=1.0)
trace = []
ws, depth, img, nears, fars, ns = OF.render_head(fo, sd, ro, rd, cf.cpu().numpy(), sd['density_bitfield'], 1, 128, sd['aabb_infer'], hp['min_near'], hp['dt_gamma'], hp['max_steps'], trace=trace)
nf = out['n_samples'].cpu().numpy()
hist = out['term_hist'].cpu().numpy()
print('oracle trace', trace)
print('hist', hist.tolist(), 'counters', out['counters'].tolist())
mism = np.nonzero(nf != ns)[0]
print('mismatch', len(mism), 'of', len(ns))
print('fused n', nf[mism][:40]); print('oracle n', ns[mism][:40])
print('ws fused', out['weights_sum'].cpu().numpy()[mism][:10], 'ws oracle', ws[mism][:10])
# loop mode on our ops
rays = utils.get_rays(fi['pose'], fi['intrinsics'], Himg, Himg)
bgc = utils.get_bg_coords(Himg, Himg, 'cuda')
with torch.no_grad():
    b = model.render(rays['rays_o'], rays['rays_d'], fi['cond'], bgc, fi['poses6'], bg_color=fi['bg_color'], reference_loop=True, loop_field='fp32', **hp)
print('loop trace', model.last_loop_trace)
d = (b['rgb_map'][0] - out['rgb_map']).abs().max(1).values.cpu().numpy()
print('fused vs loop rgb max diff', d.max(), 'n rays > 1e-5:', (d > 1e-5).sum(), 'overlap with mismatch:', np.intersect1d(np.nonzero(d > 1e-5)[0], mism).size)
from oracle import ref_gpu
if ref_gpu.available():
    ref = ref_gpu.RefRenderer(model.state_dict(), hp)
    tr = []
    with torch.no_grad():
        w2, d2, i2, n2, f2, nm = ref.render_head(rays['rays_o'][0], rays['rays_d'][0], cf, hp['dt_gamma'], hp['max_steps'], trace=tr)
    print('ref trace', tr)
    print('ref ws vs fused max', (w2 - out['weights_sum']).abs().max().item(), 'ref ws vs oracle max', np.abs(w2.cpu().numpy() - ws).max())
