"""raceid3_stemid2_package_b200 — B200-native clustering hot path of RaceID3 (compdist -> clustexp(kmedoids) ->
compmedoids) behind the reference's own API.  Compute lives in libraceid_b200.so (csrc/, C ABI in
include/raceid_b200.h); this package is the host-side mirror of the R interface."""
from .lib import Context, DistMatrix, RaceIDError, LIB_PATH, nccl_unique_id  # noqa: F401
from .scseq import (SCseq, bootstrap_samples, clustexp, clustfun, compdist, compmedoids, filterdata,  # noqa: F401
                    findoutliers_merge, outlier_distance_sample,
                    get_context, getfdata, init_distributed, refresh_partition, saturation_cln, set_context,
                    silhouette, tsne_initial_config)

__all__ = ["SCseq", "filterdata", "getfdata", "compdist", "clustexp", "clustfun", "compmedoids", "refresh_partition", "findoutliers_merge", "outlier_distance_sample", "silhouette", "tsne_initial_config",
           "Context", "DistMatrix", "RaceIDError", "get_context", "set_context", "init_distributed",
           "bootstrap_samples", "saturation_cln", "nccl_unique_id", "LIB_PATH"]
