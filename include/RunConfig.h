// include/RunConfig.h -- command line of runspECK: argv[1] = matrix, argv[2] = optional ini
// (reference include/RunConfig.h, source/RunConfig.cpp:8-23).
#pragma once
#include <string>

class RunConfig {
public:
    RunConfig(int argc, char *argv[]);
    std::string filePath;
};
