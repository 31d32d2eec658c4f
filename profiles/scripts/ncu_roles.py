"""Summarise a source-page CSV of tc_f32w_kernel by role: samples at each mbarrier wait, totals per role region."""
import csv, sys, subprocess
rep = sys.argv[1]
out = subprocess.run(['ncu','-i',rep,'--page','source','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=rows[1]
i_s=hdr.index('# Samples'); i_src=hdr.index('Source'); i_ex=hdr.index('Instructions Executed')
data=[r for r in rows[2:] if len(r)>i_s]
# duplicated listing? keep first half if duplicated
half=len(data)//2
if half and data[0][i_src]==data[half][i_src] and data[1][i_src]==data[half+1][i_src]: data=data[:half]
def I(x):
    try: return int(x)
    except: return 0
tot=sum(I(r[i_s]) for r in data)
print('kernel',rows[0][1][:60],'total samples',tot,'ninstr',len(data))
# region markers
marks=[]
for n,r in enumerate(data):
    src=r[i_src]
    for k in ('UTMALDG','UBLKCP','UTCHMMA','LDTM','STS.128','STG'):
        if k in src: marks.append((n,k))
# first occurrence of each
first={}
for n,k in marks: first.setdefault(k,n)
print('first markers',first)
print('-- instructions with >=1.2% samples or waits')
for n,r in enumerate(data):
    s=I(r[i_s])
    if s>=tot*0.012 or ('TRYWAIT' in r[i_src] and I(r[i_ex])>5000):
        print(n,s,round(100*s/tot,1),r[i_ex],r[i_src][:70])
