// Host-side entry points of the lane-parallel kernels (b2_lanes.cuh), compiled in their own translation unit
// (b2_lanes.cu) so that the engine's two halves build side by side.
#pragma once

#include <cuda_runtime.h>

#include "b2_kernels.cuh"

namespace b2 {

// Fused Panda task, G lanes per env (G chosen from the number of joints). `warps_per_block` = 0 picks the default.
template <typename T>
cudaError_t launch_task_panda_lanes(const ModelDev<T>* tables, const LaneTable<T>* lane_table, const PandaArgs<T>& a,
                                    const int* parent, const int* jtype, cudaStream_t stream, int warps_per_block = 0);

// GazeboSimulator::run of a fixed-base tree on lanes (k_run_tree_lanes): PID / force / velocity-follower joints, resets,
// external link wrenches. Not for the computed-torque controller or coupled worlds (thread kernels).
template <typename T>
cudaError_t launch_run_tree_lanes(const ModelDev<T>* tables, const LaneTable<T>* lane_table, const RunCfg<T>& cfg,
                                  const RunBuffers<T>& b, const int* parent, const int* jtype, cudaStream_t stream);

}  // namespace b2
