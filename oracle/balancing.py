"""TEST INFRASTRUCTURE ONLY.  numpy restatement of the cluster-matching half of the reference's
``estimate_balancing_weight`` (``scglue/data.py:783-836``), given cluster labels per dataset (the reference
obtains them from scanpy's leiden, ``scglue/data.py:783-791``, which is absent from this image -- see
glue_b200/data.py).  Dense arrays instead of ``sparse.COO``; the arithmetic is the reference's, line by line:

    leidens = aggregate_obs(by="leiden", obs_agg={"n": "sum"}, obsm_agg={use_rep: "mean"})   # :793-802
    us = [normalize(leiden.obsm[use_rep], norm="l2") ...]; ns = [leiden.obs["n"] ...]        # :803-804
    cosine = ui @ uj.T; cosine[cosine < cutoff] = 0; cosine = power(cosine, power)          # :809-813
    joint_cosine = prod(cosines aligned on their own axes)                                   # :814-819
    balancing = joint_cosine.sum(all axes but i) / n                                         # :826-831
    balancing = balancing.loc[leiden of every cell]; balancing /= balancing.sum() / size     # :832-835
"""
import numpy as np


def balancing_weights(reps, labels, cutoff=0.5, power=4.0):
    us, ns = [], []
    for rep, lab in zip(reps, labels):
        k = int(lab.max()) + 1
        n = np.array([(lab == c).sum() for c in range(k)], dtype=float)
        mean = np.stack([rep[lab == c].mean(axis=0) for c in range(k)])
        us.append(mean / np.linalg.norm(mean, axis=1, keepdims=True))  # sklearn normalize(norm="l2")
        ns.append(n)
    cosines = []
    for i, ui in enumerate(us):
        for j, uj in enumerate(us[i + 1:], start=i + 1):
            cosine = ui @ uj.T
            cosine[cosine < cutoff] = 0
            cosine = np.power(cosine, power)
            key = tuple(slice(None) if k in (i, j) else np.newaxis for k in range(len(us)))
            cosines.append(cosine[key])
    joint = cosines[0]
    for c in cosines[1:]:
        joint = joint * c
    out = []
    for i, (lab, n) in enumerate(zip(labels, ns)):
        balancing = joint.sum(axis=tuple(k for k in range(joint.ndim) if k != i)) / n
        w = balancing[lab]
        out.append(w / (w.sum() / w.size))
    return out
