"""CPU simulation of the conv tile order (os3d_kernel_map_order): active offsets per 128-row tile for the lexicographic
mask sort and for the frequency-ranked bit order, on the oracle's kernel maps of two synthetic frames:  python tools/sim_tile_order.py"""
import sys, numpy as np
sys.path.insert(0,'/root/repo')
from openseg3d_b200 import synthetic
from oracle import oracle
pts,_=synthetic.make_batch([0,1],1,False)
vs=[0.1,0.1,0.1]; pcr=[-72,-72,-2,72,72,4.4]
coords=np.asarray(oracle.voxelize(pts,vs,pcr)[0])
shape=[64,1440,1440]
ORDER=[13, 10,12,14,16, 9,11,15,17, 4,22, 1,3,5,7,19,21,23,25, 0,2,6,8,18,20,24,26]   # least -> most significant
def masks_of(nbr):
    mask=np.zeros(nbr.shape[0],dtype=np.uint64)
    for k in range(27): mask|=((nbr[:,k]>=0).astype(np.uint64)<<np.uint64(k))
    return mask
def permute(mask):
    pm=np.zeros(len(mask),dtype=np.uint64)
    for newpos,k in enumerate(ORDER): pm|=((mask>>np.uint64(k))&np.uint64(1))<<np.uint64(newpos)
    return pm
def act(mask, order):
    ms=mask[order]; pad=(-len(ms))%128
    ms=np.concatenate([ms,np.zeros(pad,dtype=np.uint64)]).reshape(-1,128)
    tm=np.bitwise_or.reduce(ms,axis=1)
    return np.mean([bin(int(x)).count('1') for x in tm])
cur=coords; shp=shape
for lvl in range(1,4):
    sm=oracle.strided_map(cur, shp)
    for name,nbr in (('subm',np.asarray(oracle.subm_map(cur, shp)[0])),('fwd',np.asarray(sm[2])),('inv',np.asarray(sm[3]))):
        m=nbr.shape[0]; mask=masks_of(nbr); pm=permute(mask)
        out=[]
        for br in (65536,262144):
            blk=(np.arange(m)//br).astype(np.uint64)<<np.uint64(27)
            out.append('%dk: lex %.2f perm %.2f'%(br//1024, act(mask,np.argsort(blk|mask,kind='stable')), act(mask,np.argsort(blk|pm,kind='stable'))))
        print('L%d %s rows %d  '%(lvl,name,m)+'   '.join(out))
    cur=np.asarray(sm[0]); shp=[int(x) for x in sm[1]]
