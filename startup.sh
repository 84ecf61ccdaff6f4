#!/bin/bash
# Run the ten basis experiments back to back (reference: startup.sh). Multi-GPU: prefix with
#   torchrun --standalone --local-addr 127.0.0.1 --nproc-per-node <N>
LAUNCH=${LAUNCH:-python}
nohup $LAUNCH main.py --experiments \
  configs/basis_exp/experiment_sm.yaml configs/basis_exp/experiment_mm.yaml \
  configs/basis_exp/experiment_ewc.yaml configs/basis_exp/experiment_mas.yaml \
  configs/basis_exp/experiment_icarl.yaml configs/basis_exp/experiment_fedavg.yaml \
  configs/basis_exp/experiment_fedprox.yaml configs/basis_exp/experiment_fedcurv.yaml \
  configs/basis_exp/experiment_fedweit.yaml configs/basis_exp/experiment_fedstil.yaml \
  > task.log 2>&1 &
