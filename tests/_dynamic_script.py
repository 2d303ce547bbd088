"""Replay of tests/golden/dynamic_agents.npz (a scripted sequence of Agent.create_agents / Agent.remove calls on the live
reference Model, oracle/make_golden.py::dynamic_agents_fixture) on a DeviceAgentSet; shared by the device test and the
CPU host-logic test."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def replay(device, by_mask: bool):
    from mesa_b200.agentset import DeviceAgent, DeviceAgentSet
    from mesa_b200.model import DeviceModel

    class Walker(DeviceAgent):
        pass

    g = np.load(os.path.join(GOLDEN, "dynamic_agents.npz"))
    model = DeviceModel(rng=1)
    agents = DeviceAgentSet(model, Walker, 0, {"x": torch.empty(0, dtype=torch.int64, device=device),
                                               "v": torch.empty((0, 3), dtype=torch.float32, device=device)})
    model.agents = agents
    for k in range(int(g["n_ops"])):
        if int(g[f"op{k}_kind"]) == 0:
            xs = g[f"op{k}_add_x"]
            new = agents.add_agents(len(xs), x=torch.from_numpy(xs), v=torch.from_numpy(np.repeat(xs[:, None], 3, 1) * 0.5))
            assert new.get("x").cpu().tolist() == xs.tolist()
        else:
            ids = agents.unique_ids().cpu().numpy()
            gone = np.isin(ids, g[f"op{k}_removed_ids"])
            if by_mask:
                agents.remove_agents(torch.from_numpy(gone).to(device))
            else:
                agents.remove_agents(np.nonzero(gone)[0].tolist())
        want_ids, want_x = g[f"op{k}_ids"], g[f"op{k}_x"]
        assert len(agents) == len(want_ids), k
        assert agents.unique_ids().cpu().tolist() == want_ids.tolist(), k       # survivors keep order and unique_id
        assert agents.get("x").cpu().tolist() == want_x.tolist(), k
        assert np.array_equal(agents.get("v").cpu().numpy(), np.repeat(want_x[:, None], 3, 1).astype(np.float32) * 0.5), k
        if len(agents):
            assert agents[len(agents) - 1].unique_id == int(want_ids[-1]) and agents[0].x == int(want_x[0])
    return model, agents
