import json,sys
ops=json.load(open(sys.argv[1]))
n=int(sys.argv[2]) if len(sys.argv)>2 else 40
tot=sum(o['ms'] for o in ops)
kinds={}
for o in ops:
    k=kinds.setdefault(o['kind'],[0,0,0,0]); k[0]+=o['ms']; k[1]+=o['bytes']; k[2]+=o['flops']; k[3]+=1
print("total %.3f ms over %d ops"%(tot,len(ops)))
for k,v in sorted(kinds.items(),key=lambda kv:-kv[1][0]):
    print("%-16s n=%3d %7.3f ms %5.1f%% %8.1f MB %7.1f GB/s %7.1f TF/s"%(k,v[3],v[0],100*v[0]/tot,v[1]/1e6,v[1]/v[0]/1e6,v[2]/v[0]/1e9))
print()
for o in sorted(ops,key=lambda o:-o['ms'])[:n]:
    ms=o['ms']
    print("%-16s %-42s %7.4f ms %8.1f MB %7.1f GB/s %6.1f TF/s"%(o['kind'],o['name'][:42],ms,o['bytes']/1e6,o['bytes']/ms/1e6,o['flops']/ms/1e9))
