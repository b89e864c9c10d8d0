// speck_b200/host/Config.cpp -- INI-backed run configuration (reference source/Config.cpp over
// inih): section-less `key = value` lines, ';' or '#' comments, keys case-insensitive; booleans
// accept true/yes/on/1 and false/no/off/0 like inih's GetBoolean.
#include <algorithm>
#include <cctype>
#include <cstdlib>
#include <fstream>
#include "Config.h"

namespace {
std::string trim(const std::string &s)
{
    size_t b = 0, e = s.size();
    while (b < e && std::isspace((unsigned char)s[b])) ++b;
    while (e > b && std::isspace((unsigned char)s[e - 1])) --e;
    return s.substr(b, e - b);
}
std::string lower(std::string s)
{
    std::transform(s.begin(), s.end(), s.begin(), [](unsigned char c) { return (char)std::tolower(c); });
    return s;
}
Config *g_config = nullptr;
}  // namespace

Config &Config::instance()
{
    if (!g_config) g_config = new Config();
    return *g_config;
}

const char *Config::name(Key key)
{
    switch (key) {
        case InputFile: return "inputfile";
        case IterationsWarmUp: return "iterationswarmup";
        case IterationsExecution: return "iterationsexecution";
        case TrackIndividualTimes: return "trackindividualtimes";
        case TrackCompleteTimes: return "trackcompletetimes";
        case CompareResult: return "compareresult";
        case Device: return "device";
        case Devices: return "devices";
        case GpuConvert: return "gpuconvert";
    }
    return "";
}

void Config::init()
{
    delete g_config;
    g_config = new Config();
}

void Config::init(std::string path)
{
    init();
    std::ifstream in(path);
    std::string line;
    while (std::getline(in, line)) {
        line = trim(line);
        if (line.empty() || line[0] == ';' || line[0] == '#' || line[0] == '[') continue;
        const size_t eq = line.find_first_of("=:");
        if (eq == std::string::npos) continue;
        std::string value = trim(line.substr(eq + 1));
        const size_t cmt = value.find(" ;");   // inline comments need a preceding blank, as in inih
        if (cmt != std::string::npos) value = trim(value.substr(0, cmt));
        g_config->values[lower(trim(line.substr(0, eq)))] = value;
    }
}

bool Config::lookup(Key key, std::string &out) const
{
    auto it = values.find(name(key));
    if (it == values.end()) return false;
    out = it->second;
    return true;
}

int Config::getInt(Key key, int fallback)
{
    Config &c = instance();
    auto ov = c.overrides.find((int)key);
    if (ov != c.overrides.end()) return ov->second;
    std::string s;
    if (!c.lookup(key, s)) return fallback;
    char *end = nullptr;
    const long v = std::strtol(s.c_str(), &end, 0);
    return end > s.c_str() ? (int)v : fallback;
}

int Config::setInt(Key key, int newVal) { return instance().overrides[(int)key] = newVal; }

std::string Config::getString(Key key, std::string fallback)
{
    std::string s;
    return instance().lookup(key, s) ? s : fallback;
}

bool Config::getBool(Key key, bool fallback)
{
    std::string s;
    if (!instance().lookup(key, s)) return fallback;
    s = lower(s);
    if (s == "true" || s == "yes" || s == "on" || s == "1") return true;
    if (s == "false" || s == "no" || s == "off" || s == "0") return false;
    return fallback;
}

float Config::getFloat(Key key, float fallback)
{
    std::string s;
    if (!instance().lookup(key, s)) return fallback;
    char *end = nullptr;
    const double v = std::strtod(s.c_str(), &end);
    return end > s.c_str() ? (float)v : fallback;
}
