"""The reference's proto2 schemas (protobufs/dna.proto, problem.proto, answer.proto) rebuilt at run time
with python-protobuf's descriptor API (there is no protoc in the image), so the hand-written wire
reader/writer of remy_b200/host can be cross-checked against a real protobuf implementation."""
from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

F = descriptor_pb2.FieldDescriptorProto
_TYPES = {"double": F.TYPE_DOUBLE, "uint32": F.TYPE_UINT32, "sint32": F.TYPE_SINT32, "bool": F.TYPE_BOOL}

# message -> [(number, name, type, repeated)]
DNA = {
    "Memory": [(21, "rec_send_ewma", "double", 0), (22, "rec_rec_ewma", "double", 0), (23, "rtt_ratio", "double", 0),
               (24, "slow_rec_rec_ewma", "double", 0), (25, "rtt_diff", "double", 0), (26, "queueing_delay", "double", 0)],
    "MemoryRange": [(11, "lower", "Memory", 0), (12, "upper", "Memory", 0), (13, "active_axis", "uint32", 1)],
    "Whisker": [(31, "window_increment", "sint32", 0), (32, "window_multiple", "double", 0), (33, "intersend", "double", 0),
                (34, "domain", "MemoryRange", 0)],
    "OptimizationSetting": [(41, "min_value", "double", 0), (42, "max_value", "double", 0), (43, "min_change", "double", 0),
                            (44, "max_change", "double", 0), (45, "multiplier", "double", 0), (46, "default_value", "double", 0)],
    "OptimizationSettings": [(51, "window_increment", "OptimizationSetting", 0), (52, "window_multiple", "OptimizationSetting", 0),
                             (53, "intersend", "OptimizationSetting", 0), (54, "lambda", "OptimizationSetting", 0)],
    "Range": [(61, "low", "double", 0), (62, "high", "double", 0), (63, "incr", "double", 0)],
    "ConfigRange": [(71, "link_packets_per_ms", "Range", 0), (72, "rtt", "Range", 0), (73, "num_senders", "Range", 0),
                    (74, "buffer_size", "Range", 0), (75, "mean_off_duration", "Range", 0), (76, "mean_on_duration", "Range", 0),
                    (77, "simulation_ticks", "uint32", 0), (78, "stochastic_loss_rate", "Range", 0)],
    "NetConfig": [(1, "mean_on_duration", "double", 0), (2, "mean_off_duration", "double", 0), (3, "num_senders", "uint32", 0),
                  (4, "link_ppt", "double", 0), (5, "delay", "double", 0), (6, "buffer_size", "uint32", 0),
                  (7, "stochastic_loss_rate", "double", 0), (8, "simulation_ticks", "uint32", 0)],
    "ConfigVector": [(81, "config", "NetConfig", 1)],
    "Fin": [(37, "lambda", "double", 0), (38, "domain", "MemoryRange", 0)],
    "FinTree": [(90, "domain", "MemoryRange", 0), (91, "children", "FinTree", 1), (92, "leaf", "Fin", 0),
                (93, "config", "ConfigRange", 0), (94, "optimizer", "OptimizationSettings", 0), (95, "configvector", "ConfigVector", 0)],
    "WhiskerTree": [(1, "domain", "MemoryRange", 0), (2, "children", "WhiskerTree", 1), (3, "leaf", "Whisker", 0),
                    (4, "config", "ConfigRange", 0), (5, "optimizer", "OptimizationSettings", 0), (6, "configvector", "ConfigVector", 0)],
}
PROBLEM = {
    "ProblemSettings": [(11, "prng_seed", "uint32", 0), (12, "tick_count", "uint32", 0)],
    "Problem": [(1, "settings", "ProblemSettings", 0), (2, "configs", ".RemyBuffers.NetConfig", 1), (3, "whiskers", ".RemyBuffers.WhiskerTree", 0)],
}
ANSWER = {
    "SenderResults": [(11, "throughput", "double", 0), (12, "delay", "double", 0)],
    "ThroughputsDelays": [(21, "config", ".RemyBuffers.NetConfig", 0), (22, "results", "SenderResults", 1)],
    "Outcome": [(32, "throughputs_delays", "ThroughputsDelays", 1), (33, "score", "double", 0)],
}


def _file(name, package, messages, deps=()):
    fd = descriptor_pb2.FileDescriptorProto(name=name, package=package, syntax="proto2")
    fd.dependency.extend(deps)
    for mname, fields in messages.items():
        m = fd.message_type.add(name=mname)
        for num, fname, ftype, rep in fields:
            f = m.field.add(name=fname, number=num, label=F.LABEL_REPEATED if rep else F.LABEL_OPTIONAL)
            if ftype in _TYPES:
                f.type = _TYPES[ftype]
            else:
                f.type = F.TYPE_MESSAGE
                f.type_name = ftype if ftype.startswith(".") else f".{package}.{ftype}"
    return fd


def classes():
    pool = descriptor_pool.DescriptorPool()
    pool.Add(_file("dna.proto", "RemyBuffers", DNA))
    pool.Add(_file("answer.proto", "AnswerBuffers", ANSWER, ["dna.proto"]))
    pool.Add(_file("problem.proto", "ProblemBuffers", PROBLEM, ["dna.proto"]))
    get = lambda full: message_factory.GetMessageClass(pool.FindMessageTypeByName(full))
    return {"FinTree": get("RemyBuffers.FinTree"), "WhiskerTree": get("RemyBuffers.WhiskerTree"), "ConfigRange": get("RemyBuffers.ConfigRange"),
            "NetConfig": get("RemyBuffers.NetConfig"), "Problem": get("ProblemBuffers.Problem"), "Outcome": get("AnswerBuffers.Outcome")}
