// main.cpp — command line of the B200 path tracer: same flags, short forms and defaults as the
// reference's src/main.cpp:13-24 (cxxopts there; a small parser here since cxxopts is absent):
//   -m/--mode {cpu,gpu} (gpu)  -w/--width (800)  -h/--height (450; note: -h is height, not help)
//   -s/--samples (100)  -b/--bounces (5)  -g/--gamma (2.2)  -i/--input (IronMan/IronMan.obj)
//   -o/--output (output.png)  --help
// Extras: --seed N (default 1234), --device N, --gpus N (devices 0..N-1 behind the one renderer object), --flip (write the
// image upright: the reference's PNG is upside down), --progressive N (N samples per pass, resumable accumulation),
// an --output ending in .pfm (linear float frame), --cache (binary scene cache next to the OBJ), --dump-float FILE (raw float32 W*H*3 framebuffer),
// --camera-pos x,y,z / --camera-target x,y,z / --fov deg (defaults = the constants of src/main.cpp:46-51),
// --lights x,y,z,r,g,b,I[;...] (default = the four constants of include/scene.hpp:55-80).
// Same flow as src/main.cpp:39-96: Scene -> loadFromObj -> fixed Camera -> renderer -> saveImage,
// timing uploadScene + render.  Differences, by design: --mode cpu is refused (this binary has no
// CPU renderer; the reference CPU path lives in oracle/ as test infrastructure) and a GPU failure is
// an error exit, not a silent CPU fallback (reference src/main.cpp:98-113 removed).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "b200_renderer.hpp"
#include "camera.hpp"
#include "scene.hpp"

namespace {

struct Opt { const char* longName; char shortName; bool takesValue; const char* def; const char* help; };
const Opt kOpts[] = {
    {"mode", 'm', true, "gpu", "Rendering mode (cpu/gpu)"},
    {"width", 'w', true, "800", "Image width"},
    {"height", 'h', true, "450", "Image height"},
    {"samples", 's', true, "100", "Samples per pixel"},
    {"bounces", 'b', true, "5", "Maximum ray bounces"},
    {"gamma", 'g', true, "2.2", "Gamma correction value"},
    {"input", 'i', true, "IronMan/IronMan.obj", "Input OBJ file path"},
    {"output", 'o', true, "output.png", "Output image file path"},
    {"seed", 0, true, "1234", "RNG seed (Philox key)"},
    {"device", 0, true, "0", "CUDA device ordinal"},
    {"gpus", 0, true, "1", "Render on devices 0..N-1 (scene replicated, interleaved pixel runs, one gather per frame)"},
    {"flip", 0, false, "", "Write the image upright (the reference writes it upside down)"},
    {"cache", 0, false, "", "Reuse / write the binary scene cache <input>.b2ptscene (skips OBJ parsing and BVH::build ordering)"},
    {"progressive", 0, true, "0", "Accumulate in passes of N samples per pixel (0 = one pass)"},
    {"dump-float", 0, true, "", "Also write the float framebuffer (raw float32, W*H*3)"},
    {"camera-pos", 0, true, "0,2,5", "Camera position x,y,z (reference: fixed at 0,2,5)"},
    {"camera-target", 0, true, "0,1.8,0", "Camera target x,y,z (reference: fixed at 0,1.8,0)"},
    {"fov", 0, true, "45", "Vertical field of view in degrees (reference: fixed at 45)"},
    {"lights", 0, true, "", "Replace the reference's four lights: x,y,z,r,g,b,intensity[;x,y,z,...] (at most 16)"},
    {"help", 0, false, "", "Print help"},
};

void printHelp() {
    std::cout << "GPU path tracer for NVIDIA B200 (sm_100a)\nUsage:\n  b2pt_cli [OPTION...]\n\n";
    for (const Opt& o : kOpts) {
        std::string flag = o.shortName ? std::string("  -") + o.shortName + ", --" + o.longName : std::string("      --") + o.longName;
        if (o.takesValue) flag += " arg";
        std::printf("%-28s %s%s%s%s\n", flag.c_str(), o.help, (o.def[0] ? " (default: " : ""), o.def, (o.def[0] ? ")" : ""));
    }
}

bool parseArgs(int argc, char** argv, std::map<std::string, std::string>& out, std::string& err) {
    for (const Opt& o : kOpts) if (o.takesValue) out[o.longName] = o.def;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        const Opt* opt = nullptr;
        std::string value;
        bool haveValue = false;
        if (a.rfind("--", 0) == 0) {
            std::string name = a.substr(2);
            size_t eq = name.find('=');
            if (eq != std::string::npos) { value = name.substr(eq + 1); name = name.substr(0, eq); haveValue = true; }
            for (const Opt& o : kOpts) if (name == o.longName) opt = &o;
        } else if (a.size() >= 2 && a[0] == '-') {
            for (const Opt& o : kOpts) if (o.shortName && a[1] == o.shortName) opt = &o;
            if (a.size() > 2) { value = a.substr(a[2] == '=' ? 3 : 2); haveValue = true; }
        }
        if (!opt) { err = "Option '" + a + "' does not exist"; return false; }
        if (!opt->takesValue) { out[opt->longName] = "true"; continue; }
        if (!haveValue) {
            if (i + 1 >= argc) { err = std::string("Option '") + opt->longName + "' is missing an argument"; return false; }
            value = argv[++i];
        }
        out[opt->longName] = value;
    }
    return true;
}

}  // namespace

int main(int argc, char* argv[]) {
    try {
        std::map<std::string, std::string> args;
        std::string err;
        if (!parseArgs(argc, argv, args, err)) { std::cerr << "Error: " << err << std::endl; return -1; }
        if (args.count("help")) { printHelp(); return 0; }

        const std::string mode = args["mode"], inputFile = args["input"], outputFile = args["output"];
        if (mode != "cpu" && mode != "gpu") {
            std::cerr << "Invalid rendering mode. Use 'cpu' or 'gpu'." << std::endl;   // main.cpp:114-117
            return -1;
        }
        if (mode == "cpu") {
            std::cerr << "--mode cpu is not available in this build: the engine is GPU-only and has no CPU fallback. "
                         "Use --mode gpu (the default)." << std::endl;
            return -1;
        }

        b2pt::Scene scene;
        std::cout << "Loading model from: " << inputFile << std::endl;
        if (!(args.count("cache") ? scene.loadFromObjCached(inputFile) : scene.loadFromObj(inputFile))) {
            std::cerr << "Failed to load model: " << inputFile << std::endl;   // main.cpp:40-43
            return -1;
        }
        if (!args["lights"].empty()) {
            std::vector<b2pt::Light> ls;
            const std::string& text = args["lights"];
            for (size_t pos = 0; pos <= text.size();) {
                size_t semi = text.find(';', pos);
                std::string one = text.substr(pos, semi == std::string::npos ? std::string::npos : semi - pos);
                float v[7];
                if (std::sscanf(one.c_str(), "%f,%f,%f,%f,%f,%f,%f", &v[0], &v[1], &v[2], &v[3], &v[4], &v[5], &v[6]) != 7)
                    throw std::runtime_error("Option 'lights' expects x,y,z,r,g,b,intensity[;...]");
                ls.emplace_back(b2pt::vec3(v[0], v[1], v[2]), b2pt::vec3(v[3], v[4], v[5]), v[6]);
                if (semi == std::string::npos) break;
                pos = semi + 1;
            }
            if (ls.size() > 16) throw std::runtime_error("Option 'lights': at most 16 lights");
            scene.setLights(std::move(ls));
        }
        std::cout << "- Total triangles: " << scene.getTriangles().size() << "\n- Total materials: " << scene.getMaterials().size() << std::endl;

        // main.cpp:46-51 (the defaults of the three camera options are the reference's constants)
        auto vec3Of = [](const std::string& text, const char* name) {
            float v[3];
            if (std::sscanf(text.c_str(), "%f,%f,%f", &v[0], &v[1], &v[2]) != 3)
                throw std::runtime_error(std::string("Option '") + name + "' expects x,y,z");
            return b2pt::vec3(v[0], v[1], v[2]);
        };
        b2pt::Camera camera(vec3Of(args["camera-pos"], "camera-pos"), vec3Of(args["camera-target"], "camera-target"),
                            b2pt::vec3(0.0f, 1.0f, 0.0f), static_cast<float>(std::atof(args["fov"].c_str())));

        b2pt::B200Renderer::Settings settings;
        settings.width = std::atoi(args["width"].c_str());
        settings.height = std::atoi(args["height"].c_str());
        settings.samplesPerPixel = std::atoi(args["samples"].c_str());
        settings.maxBounces = std::atoi(args["bounces"].c_str());
        settings.gamma = static_cast<float>(std::atof(args["gamma"].c_str()));

        const int gpus = std::atoi(args["gpus"].c_str());
        if (gpus < 1 || gpus > 16) throw std::runtime_error("Option 'gpus' must be between 1 and 16");
        std::vector<int> devices;
        if (gpus == 1) devices.push_back(std::atoi(args["device"].c_str()));
        else for (int d = 0; d < gpus; ++d) devices.push_back(d);
        const uint64_t seed = std::strtoull(args["seed"].c_str(), nullptr, 10);
        b2pt::B200Renderer renderer(settings, devices, seed);
        renderer.initialize();
        const int perPass = std::atoi(args["progressive"].c_str());

        auto t0 = std::chrono::high_resolution_clock::now();
        renderer.uploadScene(scene);
        if (perPass > 0) {
            renderer.renderProgressive(camera, perPass, [&](int done, const std::vector<float>&) {
                std::cout << "\rsamples per pixel: " << done << " / " << settings.samplesPerPixel << std::flush;
                return true;
            });
        } else {
            renderer.render(camera);
        }
        auto t1 = std::chrono::high_resolution_clock::now();
        double secs = std::chrono::duration<double>(t1 - t0).count();
        b2pt_stats s = renderer.stats();
        std::cout << "\nRendering completed in " << secs << " seconds" << std::endl;
        std::printf("{\"gpus\": %d, \"samples\": %lld, \"extend_rays\": %lld, \"shadow_rays\": %lld, \"fallback_rays\": %lld, \"gpu_seconds\": %.6f, "
                    "\"msamples_per_s\": %.3f, \"mrays_per_s\": %.3f}\n",
                    renderer.deviceCount(), (long long)s.samples, (long long)s.extend_rays, (long long)s.shadow_rays, (long long)s.fallback_rays, s.gpu_seconds,
                    s.samples / s.gpu_seconds * 1e-6, (s.extend_rays + s.shadow_rays) / s.gpu_seconds * 1e-6);

        renderer.saveImage(outputFile, args.count("flip") != 0);
        if (!args["dump-float"].empty()) {
            FILE* f = std::fopen(args["dump-float"].c_str(), "wb");
            if (!f) { std::cerr << "Error: cannot write " << args["dump-float"] << std::endl; return -1; }
            std::fwrite(renderer.frameBuffer().data(), sizeof(float), renderer.frameBuffer().size(), f);
            std::fclose(f);
        }
        std::cout << "Image saved as: " << outputFile << std::endl;
        return 0;
    } catch (const std::exception& e) {
        std::cerr << "Error: " << e.what() << std::endl;   // no CPU fallback
        return -1;
    }
}
