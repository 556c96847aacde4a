import sys; sys.path.insert(0,".")
from inference_gateway_b200 import SseEngine, synth, _abi as A
streams,mode=synth.make_config("C4", n_streams=4096)
eng=SseEngine(device=0,max_conns=4096,bytes_per_batch=40<<20)
slot,res=eng.process([(i,3,b) for i,(b,_,_) in enumerate(streams)])
print("recs",res.raw.n_recs,"frames",res.raw.n_frames,"decoded items",res.raw.n_decoded,"deps",res.raw.n_derived)
