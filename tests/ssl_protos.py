"""The reference's proto2 schemas (reference protos/*.proto) rebuilt at run time with python-protobuf's
descriptor API -- there is no protoc in this image.  Field numbers/types are copied from the .proto files
cited per message; used as the independent decoder/encoder for the wire-format tests."""
from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

F = descriptor_pb2.FieldDescriptorProto
_T = {"float": F.TYPE_FLOAT, "double": F.TYPE_DOUBLE, "uint32": F.TYPE_UINT32, "int32": F.TYPE_INT32, "bool": F.TYPE_BOOL}
_L = {"required": F.LABEL_REQUIRED, "optional": F.LABEL_OPTIONAL, "repeated": F.LABEL_REPEATED}


def _msg(fd, name, fields):
    m = fd.message_type.add()
    m.name = name
    for label, typ, fname, num in fields:
        f = m.field.add()
        f.name, f.number, f.label = fname, num, _L[label]
        if typ in _T:
            f.type = _T[typ]
        else:
            f.type = F.TYPE_MESSAGE
            f.type_name = "." + typ


def build():
    fd = descriptor_pb2.FileDescriptorProto()
    fd.name = "ssl_sim_all.proto"
    fd.syntax = "proto2"
    # messages_robocup_ssl_detection.proto:1-30
    _msg(fd, "SSL_DetectionBall", [("required", "float", "confidence", 1), ("optional", "uint32", "area", 2),
                                   ("required", "float", "x", 3), ("required", "float", "y", 4), ("optional", "float", "z", 5),
                                   ("required", "float", "pixel_x", 6), ("required", "float", "pixel_y", 7)])
    _msg(fd, "SSL_DetectionRobot", [("required", "float", "confidence", 1), ("optional", "uint32", "robot_id", 2),
                                    ("required", "float", "x", 3), ("required", "float", "y", 4),
                                    ("optional", "float", "orientation", 5), ("required", "float", "pixel_x", 6),
                                    ("required", "float", "pixel_y", 7), ("optional", "float", "height", 8)])
    _msg(fd, "SSL_DetectionFrame", [("required", "uint32", "frame_number", 1), ("required", "double", "t_capture", 2),
                                    ("required", "double", "t_sent", 3), ("required", "uint32", "camera_id", 4),
                                    ("repeated", "SSL_DetectionBall", "balls", 5),
                                    ("repeated", "SSL_DetectionRobot", "robots_yellow", 6),
                                    ("repeated", "SSL_DetectionRobot", "robots_blue", 7)])
    # messages_robocup_ssl_geometry.proto:1-16 (camera calibration omitted: never sent, src/serialize.cpp:78-104)
    names = ["line_width", "field_length", "field_width", "boundary_width", "referee_width", "goal_width", "goal_depth",
             "goal_wall_width", "center_circle_radius", "defense_radius", "defense_stretch", "free_kick_from_defense_dist",
             "penalty_spot_from_field_line_dist", "penalty_line_from_spot_dist"]
    _msg(fd, "SSL_GeometryFieldSize", [("required", "int32", n, i + 1) for i, n in enumerate(names)])
    _msg(fd, "SSL_GeometryData", [("required", "SSL_GeometryFieldSize", "field", 1)])
    # messages_robocup_ssl_wrapper.proto:4-7
    _msg(fd, "SSL_WrapperPacket", [("optional", "SSL_DetectionFrame", "detection", 1),
                                   ("optional", "SSL_GeometryData", "geometry", 2)])
    # grSim_Commands.proto:1-20
    _msg(fd, "grSim_Robot_Command", [("required", "uint32", "id", 1), ("required", "float", "kickspeedx", 2),
                                     ("required", "float", "kickspeedz", 3), ("required", "float", "veltangent", 4),
                                     ("required", "float", "velnormal", 5), ("required", "float", "velangular", 6),
                                     ("required", "bool", "spinner", 7), ("required", "bool", "wheelsspeed", 8),
                                     ("optional", "float", "wheel1", 9), ("optional", "float", "wheel2", 10),
                                     ("optional", "float", "wheel3", 11), ("optional", "float", "wheel4", 12)])
    _msg(fd, "grSim_Commands", [("required", "double", "timestamp", 1), ("required", "bool", "isteamyellow", 2),
                                ("repeated", "grSim_Robot_Command", "robot_commands", 3)])
    # grSim_Replacement.proto:1-19
    _msg(fd, "grSim_RobotReplacement", [("required", "double", "x", 1), ("required", "double", "y", 2),
                                        ("required", "double", "dir", 3), ("required", "uint32", "id", 4),
                                        ("required", "bool", "yellowteam", 5)])
    _msg(fd, "grSim_BallReplacement", [("required", "double", "x", 1), ("required", "double", "y", 2),
                                       ("required", "double", "vx", 3), ("required", "double", "vy", 4)])
    _msg(fd, "grSim_Replacement", [("optional", "grSim_BallReplacement", "ball", 1),
                                   ("repeated", "grSim_RobotReplacement", "robots", 2)])
    # grSim_Packet.proto:3-6
    _msg(fd, "grSim_Packet", [("optional", "grSim_Commands", "commands", 1), ("optional", "grSim_Replacement", "replacement", 2)])
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    return {m.name: message_factory.GetMessageClass(pool.FindMessageTypeByName(m.name)) for m in fd.message_type}
