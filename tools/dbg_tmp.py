import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import query_b200 as q
from gen_n1 import QUERIES, make_docs
from util_n1 import make_table
q.init(0)
docs=make_docs(3000,seed=21)
for name,where,keys,aggs in QUERIES:
    if name not in ("group_bool","group_small_int","group_string"): continue
    t=make_table(docs,where,keys,aggs); t.seal()
    qq=q.Query(t,"d",where,keys,aggs)
    print(name, qq.info)
    src=qq.kernel_source
    print([l for l in src.splitlines() if l.startswith('#define NQ_')])
    try:
        print(len(qq.execute().rows()))
    except Exception as e: print("ERR",e)
