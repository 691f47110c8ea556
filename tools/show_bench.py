import json,sys
for line in sys.stdin:
    line=line.strip()
    if not line.startswith('{'): 
        print(line[:300]); continue
    d=json.loads(line)
    print({k:d.get(k) for k in ["value","ms_per_step","parity","gpu_launches"]}, 'e2e', d['e2e'] and round(d['e2e']['value']), 'frac', round(d['roofline']['frac'],4), 'kms', round(d['roofline']['kernel_ms'],2), d['config'].get('kernel'), d['config'].get('sub_docs'))
