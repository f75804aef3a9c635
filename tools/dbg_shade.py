import sys, numpy as np, torch
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from test_shading_gpu import *
g=np.load('/root/repo/tests/golden/shading_n24.npz')
c = {k[3:]: torch.from_numpy(g[k]).cuda() for k in g.files if k.startswith("in_")}
leaves = {k: c[k].clone().requires_grad_(True) for k in ("base_color", "roughness", "viewdirs", "incidents")}
pbr, ex = run_ours(c, SoftplusLight(c["env_raw"]), leaves)
for k,ref in (("pbr",g["pbr"]),("diffuse_light",g["x_diffuse_light"]),("specular",g["x_specular"]),("incident_lights",g["x_incident_lights"]),("local_incident_lights",g["x_local_incident_lights"]),("global_incident_lights",g["x_global_incident_lights"])):
    a = npy(pbr) if k=="pbr" else npy(ex[k])
    d=np.abs(a-ref); i=np.unravel_index(d.argmax(), d.shape)
    print(k, d.max(), i, a[i], ref[i], (d>1e-5).sum())
i=np.unravel_index(np.abs(npy(ex["global_incident_lights"])-g["x_global_incident_lights"]).argmax(), g["x_global_incident_lights"].shape)
print('dir', g["in_incident_dirs"][i[0],i[1]], 'vis', g["in_visibility"][i[0],i[1]])
