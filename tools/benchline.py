import sys, json
tag = sys.argv[1] if len(sys.argv) > 1 else ""
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print(tag, "fps", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "lat_ms", round(d["latency_ms_single_stream"], 3),
      "layer_ms", round(d["roofline"]["ms_per_launch"], 4), "frac", round(d["roofline"]["frac"], 4), "nodes", d["gpu_launches_per_step"],
      d["gpu_graph_other_nodes_per_step"], d["clocks"].get("sm_mhz"))
