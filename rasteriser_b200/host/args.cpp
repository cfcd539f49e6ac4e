#include "args.hpp"

#include <cerrno>
#include <cstdlib>
#include <cstring>
#include <sstream>

namespace host {
namespace {

enum class Kind { String, UInt, Int, Float, Switch };

struct Flag {
    const char *short_name; // without dash, "" if none
    const char *long_name;  // without dashes
    Kind kind;
    bool required;
    const char *type_desc;
    const char *help;
};

// order and wording follow arguments.cpp:15-33
const Flag kFlags[] = {
    {"o", "obj", Kind::String, false, "model.obj", "path to Wavefront .obj file containing model to render"},
    {"l", "lights", Kind::String, true, "lights.csv", "path to CSV file containing directional lights in format direction_x,dir_y,dir_z,intensity,red,green,blue"},
    {"", "mats-dir", Kind::String, false, "path/", "folder containing .mtl files associated with model loaded, defaults to working directory"},
    {"x", "width", Kind::UInt, false, "pixels", "Width of output in pixels"},
    {"y", "height", Kind::UInt, false, "pixels", "Height of output in pixels"},
    {"s", "spin", Kind::Switch, false, "", "Display an animation of the model rotating around vertical (y) axis"},
    {"f", "flat", Kind::Switch, false, "", "Ignore vertex normals and use flat shading"},
    {"", "wind-clockwise", Kind::Switch, false, "", "Assume clockwise rather than anticlockwise winding angle to determine backfaces"},
    {"", "rx", Kind::Float, false, "radians", "Rotate by angle around x axis (composed as YXZ Tait-Bryan angles)."},
    {"", "ry", Kind::Float, false, "radians", "Rotate by angle around y axis (composed as YXZ Tait-Bryan angles)."},
    {"", "rz", Kind::Float, false, "radians", "Rotate by angle around z axis (composed as YXZ Tait-Bryan angles)."},
    {"", "scale", Kind::Float, false, "factor", "Scale the model by given factor"},
    {"", "dx", Kind::Float, false, "distance", "Displace model in x direction"},
    {"", "dy", Kind::Float, false, "distance", "Displace model in y direction"},
    {"", "dz", Kind::Float, false, "distance", "Displace model in z direction"},
    // extensions
    {"", "frames", Kind::UInt, false, "count", "[extension] frames in the headless spin sequence (-s), default 720"},
    {"", "save-frames", Kind::String, false, "pattern", "[extension] write every spin frame as PNG, printf pattern with one %u"},
    {"", "record", Kind::String, false, "spin.png", "[extension] write the spin sequence (-s) as one animated PNG (APNG)"},
    {"", "record-delay", Kind::UInt, false, "ms", "[extension] frame delay of --record in milliseconds (default 33)"},
    {"", "device", Kind::Int, false, "index", "[extension] CUDA device to render on"},
    {"", "frame-out", Kind::String, false, "file.png", "[extension] colour output (default frame.png)"},
    {"", "depth-out", Kind::String, false, "file.png", "[extension] depth output (default depth.png)"},
    {"", "timing", Kind::Switch, false, "", "[extension] print the wall time of each stage (load, context, upload + draw, outputs) to stderr"},
    {"", "quiet", Kind::Switch, false, "", "[extension] no progress output"},
    {"", "mesh-cache", Kind::String, false, "file.rastmesh", "[extension] binary copy of the parsed model: read it if present, else parse the .obj and write it"},
    {"", "load-threads", Kind::UInt, false, "count", "[extension] threads parsing the .obj (default: all hardware threads)"},
    {"", "flat-mode", Kind::String, false, "reference|face", "[extension] what -f does: 'reference' = nothing, like the reference (default); 'face' = one normal per face"},
    {"", "material-mode", Kind::String, false, "reference|kd-texture", "[extension] textured materials: 'reference' = the texture replaces Kd, like the reference (default); 'kd-texture' = texel x Kd"},
};
const int kNumFlags = (int)(sizeof kFlags / sizeof kFlags[0]);

const Flag *find_flag(const std::string &arg) {
    if (arg.size() >= 3 && arg[0] == '-' && arg[1] == '-') {
        for (const Flag &f : kFlags)
            if (arg.compare(2, std::string::npos, f.long_name) == 0) return &f;
    } else if (arg.size() == 2 && arg[0] == '-') {
        for (const Flag &f : kFlags)
            if (f.short_name[0] && arg[1] == f.short_name[0]) return &f;
    }
    return nullptr;
}

std::string flag_id(const Flag &f) {
    std::string s;
    if (f.short_name[0]) s = std::string("-") + f.short_name + " (--" + f.long_name + ")";
    else s = std::string("(--") + f.long_name + ")";
    return s;
}

} // namespace

std::string usage_text(const char *program) {
    std::ostringstream os;
    os << "USAGE: \n\n   " << program << "  -l <lights.csv> [-o <model.obj>] [--mats-dir <path/>] [-x <pixels>] [-y <pixels>] [-s] [-f]\n"
       << "        [--wind-clockwise] [--rx <radians>] [--ry <radians>] [--rz <radians>] [--scale <factor>]\n"
       << "        [--dx <distance>] [--dy <distance>] [--dz <distance>] [--] [--version] [-h]\n\nWhere: \n\n";
    for (const Flag &f : kFlags) {
        os << "   ";
        if (f.short_name[0]) os << "-" << f.short_name << (f.kind == Kind::Switch ? "" : std::string(" <") + f.type_desc + ">") << ",  ";
        os << "--" << f.long_name << (f.kind == Kind::Switch ? "" : std::string(" <") + f.type_desc + ">") << "\n     " << (f.required ? "(required)  " : "") << f.help << "\n\n";
    }
    os << "   --,  --ignore_rest\n     Ignores the rest of the labeled arguments following this flag.\n\n"
       << "   --version\n     Displays version information and exits.\n\n   -h,  --help\n     Displays usage information and exits.\n\n"
       << "   Render a model by rasterisation.\n";
    return os.str();
}

ParseResult parse_args(int argc, const char *const *argv, Args &args, std::string &message) {
    bool seen[64] = {false};
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        if (a == "--" || a == "--ignore_rest") break;
        if (a == "-h" || a == "--help") return ParseResult::Help;
        if (a == "--version") return ParseResult::Version;
        const Flag *f = find_flag(a);
        if (!f) { message = "PARSE ERROR: Argument: " + a + "\n             Couldn't find match for argument"; return ParseResult::Error; }
        const int idx = (int)(f - kFlags);
        if (seen[idx]) { message = "PARSE ERROR: Argument: " + flag_id(*f) + "\n             Argument already set!"; return ParseResult::Error; }
        seen[idx] = true;
        std::string value;
        if (f->kind != Kind::Switch) {
            if (i + 1 >= argc) { message = "PARSE ERROR: Argument: " + flag_id(*f) + "\n             Missing a value for this argument!"; return ParseResult::Error; }
            value = argv[++i];
        }
        char *end = nullptr;
        errno = 0;
        double num = 0;
        if (f->kind == Kind::UInt || f->kind == Kind::Int || f->kind == Kind::Float) {
            num = (f->kind == Kind::Float) ? std::strtod(value.c_str(), &end) : (double)std::strtoll(value.c_str(), &end, 10);
            const bool negative_unsigned = f->kind == Kind::UInt && value.find('-') != std::string::npos;
            if (end == value.c_str() || *end != '\0' || errno != 0 || negative_unsigned) {
                message = "PARSE ERROR: Argument: " + flag_id(*f) + "\n             Couldn't read argument value from string '" + value + "'";
                return ParseResult::Error;
            }
        }
        const std::string n = f->long_name;
        if (n == "obj") args.obj_file = value;
        else if (n == "lights") args.lights_file = value;
        else if (n == "mats-dir") args.materials_directory = value;
        else if (n == "width") args.image_width = (unsigned)num;
        else if (n == "height") args.image_height = (unsigned)num;
        else if (n == "spin") args.spin = true;
        else if (n == "flat") args.flat = true;
        else if (n == "wind-clockwise") args.wind_clockwise = true;
        else if (n == "rx") args.tait_bryan_angles[0] = (float)num;
        else if (n == "ry") args.tait_bryan_angles[1] = (float)num;
        else if (n == "rz") args.tait_bryan_angles[2] = (float)num;
        else if (n == "scale") args.scale = (float)num;
        else if (n == "dx") args.displacement[0] = (float)num;
        else if (n == "dy") args.displacement[1] = (float)num;
        else if (n == "dz") args.displacement[2] = (float)num;
        else if (n == "frames") args.frames = (unsigned)num;
        else if (n == "save-frames") args.save_frames = value;
        else if (n == "record") args.record = value;
        else if (n == "record-delay") args.record_delay_ms = (unsigned)num;
        else if (n == "device") args.device = (int)num;
        else if (n == "frame-out") args.frame_out = value;
        else if (n == "depth-out") args.depth_out = value;
        else if (n == "quiet") args.quiet = true;
        else if (n == "timing") args.timing = true;
        else if (n == "mesh-cache") args.mesh_cache = value;
        else if (n == "load-threads") args.load_threads = (unsigned)num;
        else if (n == "flat-mode") {
            if (value != "reference" && value != "face") { message = "PARSE ERROR: Argument: (--flat-mode)\n             Value '" + value + "' does not meet constraint: reference|face"; return ParseResult::Error; }
            args.flat_face = value == "face";
        }
        else if (n == "material-mode") {
            if (value != "reference" && value != "kd-texture") { message = "PARSE ERROR: Argument: (--material-mode)\n             Value '" + value + "' does not meet constraint: reference|kd-texture"; return ParseResult::Error; }
            args.modulate_kd = value == "kd-texture";
        }
    }
    for (int k = 0; k < kNumFlags; ++k)
        if (kFlags[k].required && !seen[k]) {
            message = "PARSE ERROR: \n             Required argument missing: " + std::string(kFlags[k].long_name);
            return ParseResult::Error;
        }
    args.aspect_ratio = (float)args.image_width / (float)args.image_height; // arguments.cpp:39
    return ParseResult::Ok;
}

} // namespace host
