import sys, time; sys.path.insert(0,'/root/repo')
import numpy as np
from oracle import oracle
from subrecon_b200 import treeio, synth, native
def cmp(name, r, lnl, joint):
    el=np.max(np.abs(r['lnl']-lnl)/np.abs(lnl))
    big=joint>1e-300
    ej=np.max(np.abs(r['joint'][big]-joint[big])/joint[big])
    ea=np.max(np.abs(r['joint']-joint))
    print(f"{name}: max rel lnL {el:.3e}  max rel joint {ej:.3e}  max abs joint {ea:.3e}  sum {r['joint'].sum(axis=(1,2)).min():.15f}", flush=True)
ctx=native.Context(0)
print("fp64 peak", ctx.measure_fp64_peak(100))
# config 1
import os
ex='/root/repo/tests/golden/example/'
t=treeio.read_tree(ex+'root.tre'); names,seqs=treeio.read_fasta(ex+'prot.fasta'); st=treeio.encode_alignment(seqs)
ft=treeio.flatten_tree(t,{n:i for i,n in enumerate(names)})
freqs=np.array([float(x) for x in open(ex+'freqs').read().split(',')]); freqs/=freqs.sum()
m=oracle.OracleModel('WAG',freqs); rates=oracle.gamma_rates(4,0.5)
lnl,joint=oracle.run(ft,st,m,rates)
for tm in (2,4):
    ctx.set_option('tile_m',tm)
    ctx.set_model(m.Eval,m.Evec,m.Ievc,m.pi); ctx.set_rates(rates); ctx.set_tree(ft); ctx.set_alignment(st)
    print(ctx.schedule_info())
    r=ctx.run(); cmp(f"lysozyme tile_m={tm}", r, lnl, joint)
    P=ctx.debug_pmatrix(5,1); Po=oracle.transition_probs(m.Eval,m.Evec,m.Ievc,ft.blen[5]*rates[1]); print("  P rel", np.max(np.abs(P-Po)/Po))
    r2=ctx.run_streamed(st); cmp("  streamed", r2, lnl, joint)
    r3=ctx.run(site0=37,n=5); print("  single-site window", np.max(np.abs(r3['lnl']-lnl[37:42])))
# config 2
m=oracle.OracleModel('JTT'); rates=oracle.gamma_rates(4,0.5)
from subrecon_b200 import palmodel
hm=palmodel.host_model('jtt')
t=synth.balanced_tree(100,2); st=synth.simulate_alignment(t,hm,rates,10000,102); ft=treeio.flatten_tree(t)
t0=time.time(); lnl,joint=oracle.run(ft,st,m,rates,threads=16); print("oracle", time.time()-t0)
for tm in (2,4):
    ctx.set_option('tile_m',tm)
    ctx.set_model(m.Eval,m.Evec,m.Ievc,m.pi); ctx.set_rates(rates); ctx.set_tree(ft); ctx.set_alignment(st)
    print(ctx.schedule_info())
    t0=time.time(); r=ctx.run(); dt=time.time()-t0; cmp(f"cfg2 tile_m={tm} ({dt*1e3:.1f} ms)", r, lnl, joint)
# config 3 subset + timing
m=oracle.OracleModel('WAG'); hm=palmodel.host_model('wag')
t=synth.yule_tree(1000,3); ft=treeio.flatten_tree(t)
st=synth.simulate_alignment(t,hm,rates,4096,103)
t0=time.time(); lnl,joint=oracle.run(ft,st,m,rates,threads=16); print("oracle 4096 cols", time.time()-t0)
big=np.tile(st,(1,64))   # 262144 columns
for tm in (2,4):
    ctx.set_option('tile_m',tm); ctx.set_option('profile',1)
    ctx.set_model(m.Eval,m.Evec,m.Ievc,m.pi); ctx.set_rates(rates); ctx.set_tree(ft); ctx.set_alignment(st)
    info=ctx.schedule_info(); print(info)
    r=ctx.run(); cmp(f"cfg3 tile_m={tm}", r, lnl, joint)
    ctx.set_alignment(big)
    ctx.run(want_joint=False); ctx.profile(reset=True)
    for i in range(3): ctx.run(want_joint=False)
    p=ctx.profile(reset=True); print(p)
    cols=p['columns_pruned']; ms=p['ms_prune']
    print(f"  K2: {cols/ms*1e3/1e6:.3f} M col/s, {cols*4*info['flops_per_column_rate']/ms*1e3/1e12:.2f} TFLOP/s algorithmic; K3 {p['ms_root']/p['n_root']:.3f} ms/launch")
