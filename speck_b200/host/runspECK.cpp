// speck_b200/host/runspECK.cpp -- the CLI: runspECK <matrix.mtx> [config.ini]
// (reference source/runspECK.cpp:13-32).  fp64 unless argv[7] == "f" (the reference parses that
// argument and then ignores it).
#include <cstdio>
#include <exception>
#include <string>
#include "Executor.h"

int main(int argc, char *argv[])
{
    if (argc < 2) {
        printf("no .mtx file path set. please call using 'runspECK /path/to/matrix.mtx [config.ini]'\n");
        return -1;
    }
    const std::string valueType = argc > 7 ? argv[7] : "d";
    try {
        if (valueType == "f") return Executor<float>(argc, argv).run();
        return Executor<double>(argc, argv).run();
    } catch (const char *msg) {
        printf("%s\n", msg);
        return 1;
    } catch (std::exception &e) {
        printf("%s\n", e.what());
        return 1;
    }
}
