// include/Config.h -- key/value run configuration read from an INI file (reference
// include/Config.h + source/Config.cpp over inih).  Only the six keys the reference driver ever
// reads are kept, plus `Device` for selecting the GPU and `Devices` (comma-separated list,
// e.g. Devices=0,1,2,3) for the row-sharded multi-GPU run and `GpuConvert` (COO -> CSR of a freshly
// parsed .mtx on the device instead of the host sort).
#pragma once
#include <map>
#include <string>

class Config {
public:
    enum Key { InputFile, IterationsWarmUp, IterationsExecution, TrackIndividualTimes, TrackCompleteTimes,
               CompareResult, Device, Devices, GpuConvert };

    static void init(std::string path);   // parse `path` (section-less key=value, ';'/'#' comments)
    static void init();                   // no file: every get* returns its fallback
    static int getInt(Key key, int fallback = -1);
    static int setInt(Key key, int newVal);
    static std::string getString(Key key, std::string fallback = "");
    static bool getBool(Key key, bool fallback = false);
    static float getFloat(Key key, float fallback = 0.0f);

private:
    static Config &instance();
    static const char *name(Key key);
    bool lookup(Key key, std::string &out) const;
    std::map<std::string, std::string> values;   // lower-cased key -> raw value
    std::map<int, int> overrides;
};
