// TEST INFRASTRUCTURE ONLY. Empty stand-in: the reference includes <THC/THC.h>
// (lib/pointgroup_ops/src/bfs_cluster/bfs_cluster.h:11), a header removed from torch >= 1.11;
// nothing from it is used.
#pragma once
