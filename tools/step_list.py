import csv,re,sys
rows=list(csv.reader(open(sys.argv[1])))
hi=[i for i,r in enumerate(rows) if "Kernel Name" in r][0]
h=rows[hi]; kn=h.index("Kernel Name"); mv=h.index("Metric Value")
data=[(re.sub(r"\(.*","",r[kn]).replace("void ",""),float(r[mv].replace(",",""))/1e3) for r in rows[hi+1:] if len(r)>mv and r[mv]]
idx=[i for i,(k,v) in enumerate(data) if k.startswith("knn_bbox_init")]
step=data[idx[-2]:idx[-1]]
print(len(step), sum(v for k,v in step))
for k,v in step:
    if "at::" in k or "loss" in k or "Cat" in k or "cat" in k: print(round(v,1),k[:80])
