import sys, numpy as np
sys.path.insert(0,'.')
from emplanner_carla_b200 import api, scenes
for name in ('C','A'):
    lat=scenes.LATTICES[name]
    sc=scenes.sl_scenes(4096,lat,{'A':0,'C':2}[name])
    dp,qp=api.dp_params(**lat.kwargs()),api.qp_params()
    pl=api.HostPlanner(4096,8,api.max_enriched(dp),lat.col)
    v,_=pl.plan(api.pack_scenes(**sc),dp,qp)
    st=v.qp_status.astype(int); it=v.qp_iters.astype(int)
    for s in np.unique(st):
        m=st==s
        print(name,'status',s,'count',m.sum(),'iters: mean %.1f p50 %d p90 %d p99 %d max %d'%(it[m].mean(),np.percentile(it[m],50),np.percentile(it[m],90),np.percentile(it[m],99),it[m].max()), 'gap median %.2e'%np.median(v.gap[m]))
    print(name,'hist optimal',np.bincount(it[st==0])[:25])
    print(name,'hist unknown',np.bincount(it[st!=0])[:45])
    pl.close()
