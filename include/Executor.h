// include/Executor.h -- benchmark driver behind runspECK (reference include/Executor.h,
// source/Executor.cpp:13-81).
#pragma once
#include "RunConfig.h"

template <typename ValueType>
class Executor {
public:
    Executor(int argc, char *argv[]) : runConfig(argc, argv) {}
    int run();

private:
    RunConfig runConfig;
    int iterationsWarmup = 0;
    int iterationsExecution = 1;
};
