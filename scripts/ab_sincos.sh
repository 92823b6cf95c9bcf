for v in fast lib fast lib; do
  if [ $v = lib ]; then export B2SIM_LIBRARY=$PWD/gym-ignition_b200/lib/libb2sim_libsincos.so; else unset B2SIM_LIBRARY; fi
  echo "== $v"
  python scripts/traj_probe.py 2>&1 | grep "Pendulum-Gazebo-v0 n=65536\|Swingup-Gazebo-v0 n=1048576"
  B2_CHAIN_KERNEL=plain python scripts/chain_probe.py --child 4194304 2>&1 | head -1
  B2_PANDA_KERNEL=thread python scripts/panda_lanes_probe.py --child 16384 float64 gpurun_out/x.npz 2>&1 | tail -1
  B2_PANDA_KERNEL=lanes python scripts/panda_lanes_probe.py --child 16384 float64 gpurun_out/x.npz 2>&1 | tail -1
done
