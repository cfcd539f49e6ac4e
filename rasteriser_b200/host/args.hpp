// args.hpp -- command line of the host program: the reference's flag table (arguments.cpp:12-48,
// struct Args in headers/arguments.h:7-21) parsed by hand (TCLAP is not available offline), plus a few
// extension flags that do not collide with it.
#pragma once

#include <string>

namespace host {

struct Args {
    // reference fields (headers/arguments.h:8-20) with the defaults of arguments.cpp:15-33
    unsigned int image_width = 540u;
    unsigned int image_height = 304u;
    float aspect_ratio = 540.f / 304.f;
    bool spin = false;
    bool flat = false; // parsed and carried; like the reference, the frame path never reads it
    bool wind_clockwise = false;
    float scale = 1.f;
    float displacement[3] = {0.f, 0.f, 0.f};
    float tait_bryan_angles[3] = {0.f, 0.f, 0.f}; // rx, ry, rz
    std::string obj_file;
    std::string lights_file;
    std::string materials_directory;

    // extensions (not in the reference)
    unsigned int frames = 720;        // --frames N: length of the headless spin sequence
    std::string save_frames;          // --save-frames PATTERN: printf pattern with one %u, e.g. spin_%04u.png
    std::string record;               // --record FILE: the spin sequence as one animated PNG
    unsigned int record_delay_ms = 33; // --record-delay MS
    int device = 0;                   // --device N
    std::string frame_out = "frame.png", depth_out = "depth.png"; // --frame-out / --depth-out
    bool quiet = false;               // --quiet
    bool timing = false;              // --timing: stage wall times on stderr
    std::string mesh_cache;           // --mesh-cache FILE: binary copy of the parsed model (read if present, else written after parsing)
    unsigned int load_threads = 0;    // --load-threads N: OBJ parser threads (0 = all hardware threads)
    bool modulate_kd = false;         // --material-mode kd-texture: texel x Kd for textured materials (extension; the reference drops Kd there)
    bool flat_face = false;           // --flat-mode face: with -f, shade with one normal per face (extension; default keeps the reference's no-op)
};

enum class ParseResult { Ok, Help, Version, Error };

// Fills `args`; on Error `message` holds the text the program prints before exit(1)
// (TCLAP prints "PARSE ERROR: ..." plus a usage hint; arguments.cpp:49-50).
ParseResult parse_args(int argc, const char *const *argv, Args &args, std::string &message);
std::string usage_text(const char *program);

} // namespace host
