// Registry of the games compiled into the engine. PG2_FOR_EACH_GAME(X) expands X(name, Type).
#pragma once
#include "bossfight.cuh"
#include "caveflyer.cuh"
#include "chaser.cuh"
#include "climber.cuh"
#include "coinrun.cuh"
#include "jumper.cuh"
#include "maze.cuh"

#define PG2_FOR_EACH_GAME(X) \
    X("maze", pg2::Maze)         \
    X("coinrun", pg2::CoinRun)   \
    X("bossfight", pg2::BossFight) \
    X("climber", pg2::Climber)     \
    X("chaser", pg2::Chaser)       \
    X("caveflyer", pg2::CaveFlyer) \
    X("jumper", pg2::Jumper)

// Distribution modes that change the world size are separate instantiations (no cost for the default mode):
// PG2_FOR_EACH_GAME_MODE(X) expands X(name, mode, Type) for every non-default (game, mode) pair that is built.
#define PG2_FOR_EACH_GAME_MODE(X) \
    X("maze", 0, pg2::MazeT<0>)    \
    X("maze", 2, pg2::MazeT<2>)    \
    X("chaser", 1, pg2::ChaserT<1>) \
    X("chaser", 2, pg2::ChaserT<2>) \
    X("jumper", 0, pg2::JumperT<0>)   \
    X("jumper", 2, pg2::JumperT<2>)   \
    X("caveflyer", 0, pg2::CaveFlyerT<0>) \
    X("caveflyer", 2, pg2::CaveFlyerT<2>) \
    X("bossfight", 0, pg2::BossFightT<0>)
