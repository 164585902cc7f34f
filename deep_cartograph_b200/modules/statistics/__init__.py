from . import statistics  # noqa: F401
from .statistics import cluster_data, kmeans_clustering, find_centroids, optimize_clustering, cluster_scores  # noqa: F401
