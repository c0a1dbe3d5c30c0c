import csv,sys,subprocess
rep=sys.argv[1]
out=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=rows[0]; units=rows[1]; vals=rows[2]
keys=sys.argv[2:] 
for i,h in enumerate(hdr):
    if any(k in h for k in keys): print(h, '=', vals[i], units[i])
