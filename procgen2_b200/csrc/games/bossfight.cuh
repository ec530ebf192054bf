// BossFight — device restatement of /root/reference/games/bossfight/:
//   step logic  cenv_step bossfight.cpp:294-347; System_Agent::update common_systems.cpp:494-683;
//               System_Mob_AI::update :199-390 (fire :75, explode :90, fire_pattern :103-185, show_damage :187)
//   level gen   reset() bossfight.cpp:426-504; System_Agent::reset :723-737; System_Mob_AI::reset :452-469
//   frame       render_game bossfight.cpp:401-424; System_Mob_AI::render :392-450; System_Agent::render :685-721;
//               System_Sprite_Render::render :25-49
// hard_mode (compile-time default, common_systems.h:61); easy_mode is the BossFightT<0> instantiation. World = the 64x64 px observation at
// camera scale 1 => screen rectangle {-2,-2,4,4} units (obs-only rendering).
// Entity ids per episode: 0 player, 1 boss, 2.. accepted barriers.
#pragma once
#include "../pg2_common.cuh"
#include "../pg2_libm.cuh"
#include "../pg2_render.cuh"
#include "../pg2_rng.cuh"
#include "../pg2_state.cuh"
#include "../pg2_uset.cuh"
#include "../pg2_warp.cuh"

namespace pg2 {

// Pools are slot-major: field[slot * N + env].
#define PG2_BOSSFIGHT_FIELDS(F)                                                                 \
    F(float, px, 1) F(float, py, 1) F(float, pvx, 1) F(float, pvy, 1)   /* player */            \
    F(float, bx, 1) F(float, by, 1) F(float, bvx, 1) F(float, bvy, 1)   /* boss */              \
    F(float, phase_timer, 1) F(int32_t, phase_index, 1) F(int32_t, weapon_index, 1)            \
    F(float, attack_timer, 1) F(int32_t, hp, 1)                                                 \
    F(int32_t, m_next_bullet, 1) F(int32_t, m_next_expl, 1)                                     \
    F(int32_t, m_num_bullets, 1) F(int32_t, m_num_expl, 1)                                      \
    F(float, expl_timer, 1) F(float, damage_timer, 1) F(float, move_timer, 1)                   \
    F(int32_t, m_ship, 1) F(int32_t, m_bullet_tex, 1)                                           \
    F(float, mb_x, 64) F(float, mb_y, 64) F(float, mb_vx, 64) F(float, mb_vy, 64)               \
    F(float, mb_rot, 64) F(float, mb_frame, 64)                                                 \
    F(float, ex_x, 8) F(float, ex_y, 8) F(float, ex_frame, 8)                                   \
    F(int32_t, a_next_bullet, 1) F(int32_t, a_num_bullets, 1) F(float, a_bullet_timer, 1)       \
    F(int32_t, a_ship, 1) F(int32_t, a_bullet_tex, 1) F(uint8_t, alive, 1)                      \
    F(float, ab_x, 32) F(float, ab_y, 32) F(float, ab_vx, 32) F(float, ab_vy, 32)               \
    F(float, ab_frame, 32) F(uint8_t, ab_bouncing, 32) F(float, ab_bounce_timer, 32)            \
    F(int32_t, num_barriers, 1) F(float, bar_x, 4) F(float, bar_y, 4) F(uint8_t, bar_tex, 4)    \
    F(int32_t, n_hazards, 1) F(uint8_t, hazard_order, 8)   /* iteration order of System_Hazard::entities (ids) */ \
    F(uint8_t, sprite_order, 4)                            /* ... of System_Sprite_Render::entities (ids) */      \
    F(int32_t, nb_hazard, 1) F(int32_t, nb_sprite, 1)      /* persisted bucket counts (Q25) */  \
    F(int32_t, bg_index, 1)

PG2_DEFINE_STATE(BossFightState, PG2_BOSSFIGHT_FIELDS)

template <int MODE>
struct BossFightT {
    using State = BossFightState;
    static constexpr int SUB_STEPS = 4;
    static constexpr bool LANE_AWARE = true;    // step(): the bullet rings are walked lane-parallel, the rest is uniform (leader stores)
    static constexpr int STEP_LANES = 8;        // lanes per environment in k_step: four environments share a warp (measured: 4 / 16 / 32 lanes are 5-15 % slower)
    static constexpr int MAX_POST = 112;        // capacity of the frame's post-blit list
    static constexpr bool ROTATES = true;     // some blits are rotated
    static constexpr bool SLOW_RESET = false;   // level generation is long: run it concurrently with the render of the other envs
    static constexpr int RESET_ARENA = 2 * 1024;   // per-warp level-generation scratch (high water measured with PG2_ARENA_TRACE)
    static constexpr bool PREFETCH_LEVELS = false;   // step() draws from the RNG: the next level is not known ahead of time
    static constexpr int PREFETCH_MIN_EPISODE = 0;   // level prefetch whatever max_episode_steps is
    static const char* reset_keeps() { return ""; }
    static constexpr int TILE_CLASSES = 1;
    static constexpr int WIN_ROWS = 1;        // most tile rows the camera window can span (zoom-dependent; frame table sizing)
    static constexpr int BLIT_UNROLL = 2;     // post-blit patches fetched together (pg2_render.cuh draw_blit_band)
    static constexpr int RENDER_MIN_CTAS = 10;   // CTAs per SM the register allocation of k_render aims at (small frame tables: 10 frames per SM measured +5 % over 8)
    static constexpr int DEFAULT_MODE = MODE;    // this instantiation's distribution mode (the reference compiles in 1 = hard; common_systems.h:61-65)
    static bool mode_supported(int mode) { return mode == MODE; }
    // System_Mob_AI::Config::mode: easy = boss bullets at half speed, shorter shielded phases. Compile-time: as run-time
    // selects the two constants cost k_step 38 registers (118 -> 156) and with them a quarter of its occupancy.
    static constexpr float BULLET_SPEED = MODE == 0 ? 0.05f : 0.1f;     // common_systems.cpp:104
    static constexpr float SHIELD_JITTER = MODE == 0 ? 30.0f : 80.0f;   // common_systems.cpp:202
    static constexpr bool HAS_TILES = false;     // the frame has a tile layer
    static constexpr bool STATIC_VIEW = true;    // fixed camera, no tile layer: the background image is cached per env
    static constexpr int MB = 64, NEX = 8, AB = 32, MAX_BAR = 4;
    enum Tex {
        T_BOSS0 = 0,       // 4 enemy ships
        T_BULLET0 = 4,     // 3 lasers
        T_EXPL0 = 7,       // 5 explosion frames
        T_SHIELD = 12,
        T_PLAYER0 = 13,    // 4 player ships
        T_BARRIER0 = 17,   // 8 meteors
        T_BG0 = 25, NUM_BG = 13,
        NUM_TEX = 38
    };

    static const char* const* texture_names(int* count) {
        static const char* const names[NUM_TEX] = {
            "assets/misc_assets/enemyShipBlack1.png", "assets/misc_assets/enemyShipBlue2.png",
            "assets/misc_assets/enemyShipGreen3.png", "assets/misc_assets/enemyShipRed4.png",
            "assets/misc_assets/laserGreen14.png", "assets/misc_assets/laserRed11.png", "assets/misc_assets/laserBlue09.png",
            "assets/misc_assets/explosion1.png", "assets/misc_assets/explosion2.png", "assets/misc_assets/explosion3.png",
            "assets/misc_assets/explosion4.png", "assets/misc_assets/explosion5.png",
            "assets/misc_assets/shield2.png",
            "assets/misc_assets/playerShip1_blue.png", "assets/misc_assets/playerShip1_green.png",
            "assets/misc_assets/playerShip2_orange.png", "assets/misc_assets/playerShip3_red.png",
            "assets/misc_assets/spaceMeteors_001.png", "assets/misc_assets/spaceMeteors_002.png",
            "assets/misc_assets/spaceMeteors_003.png", "assets/misc_assets/spaceMeteors_004.png",
            "assets/misc_assets/meteorGrey_big1.png", "assets/misc_assets/meteorGrey_big2.png",
            "assets/misc_assets/meteorGrey_big3.png", "assets/misc_assets/meteorGrey_big4.png",
            "assets/space_backgrounds/deep_space_01.png", "assets/space_backgrounds/spacegen_01.png",
            "assets/space_backgrounds/milky_way_01.png", "assets/space_backgrounds/ez_space_lite_01.png",
            "assets/space_backgrounds/meyespace_v1_01.png", "assets/space_backgrounds/eye_nebula_01.png",
            "assets/space_backgrounds/deep_sky_01.png", "assets/space_backgrounds/space_nebula_01.png",
            "assets/space_backgrounds/Background-1.png", "assets/space_backgrounds/Background-2.png",
            "assets/space_backgrounds/Background-3.png", "assets/space_backgrounds/Background-4.png",
            "assets/space_backgrounds/parallax-space-backgound.png",
        };
        *count = NUM_TEX;
        return names;
    }

    // ---------------------------------------------------------------------------------------
    // System_Mob_AI::fire (common_systems.cpp:75-88)
    struct MobPool {   // every lane keeps the (identical) counters, the leader stores
        const State& s; int env, N;
        int next_bullet, num_bullets, next_expl, num_expl;
        bool leader;
        PG2_DEV void fire(float x, float y, float rotation, float speed) {
            if (num_bullets < MB) {
                int i = next_bullet * N + env;
                float sn, cs;
                glibc_sincosf(rotation, &sn, &cs);
                if (leader) {
                    s.mb_rot[i] = rotation;
                    s.mb_vx[i] = __fmul_rn(cs, speed);
                    s.mb_vy[i] = __fmul_rn(-sn, speed);
                    s.mb_x[i] = x; s.mb_y[i] = y;
                    s.mb_frame[i] = 0.0f;
                }
                next_bullet = (next_bullet + 1) % MB;
                num_bullets++;
            }
        }
        PG2_DEV void explode(float x, float y) {
            if (num_expl < NEX) {
                int i = next_expl * N + env;
                if (leader) { s.ex_x[i] = x; s.ex_y[i] = y; s.ex_frame[i] = 0.0f; }
                next_expl = (next_expl + 1) % NEX;
                num_expl++;
            }
        }
    };

    // A bullet ring is walked newest to oldest, `i < num` with num shrinking by one whenever a bullet is destroyed
    // (common_systems.cpp:330-372, 560-640): entry i is visited iff i + #destroyed among the entries before it < n0 —
    // a prefix of the ring, since that sum grows with i. D: bit i = entry i would be destroyed if visited (an entry's own
    // fate does not depend on the others). Returns the number of visited entries.
    static PG2_DEV int visited_prefix(uint64_t D, int n0) {
        if ((D & (n0 >= 64 ? ~0ull : (1ull << n0) - 1ull)) == 0ull) return n0;   // nothing destroyed: the whole ring
        int lo = 0, hi = n0;   // the largest L <= n0 with (L - 1) + popc(D below L - 1) < n0, by bisection (monotone)
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1, i = mid - 1;
            const int g = i + __popcll(D & ((1ull << i) - 1ull));
            if (g < n0) lo = mid; else hi = mid - 1;
        }
        return lo;
    }

    static PG2_DEV_NOINLINE bool step(const State& s, const CommonState& c, int env, int action, float* reward, const StepCtx& ctx) {
        const int N = s.N;
        const float dt = 1.0f / SUB_STEPS;
        const double PI = 3.14159265358979323846;
        Mt rng; rng.mt = c.mt + (size_t)env * MT_N; rng.idx = c.mti[env];
        rng.collective = ctx.nlanes != 1; rng.writer = ctx.leader(); rng.sync_mask = ctx.mask;
        const Rect screen{ -2.0f, -2.0f, 4.0f, 4.0f };

        float px = s.px[env], py = s.py[env], pvx = s.pvx[env], pvy = s.pvy[env];
        float bx = s.bx[env], by = s.by[env], bvx = s.bvx[env], bvy = s.bvy[env];
        float phase_timer = s.phase_timer[env], attack_timer = s.attack_timer[env];
        int phase_index = s.phase_index[env], weapon_index = s.weapon_index[env], hp = s.hp[env];
        float expl_timer = s.expl_timer[env], damage_timer = s.damage_timer[env], move_timer = s.move_timer[env];
        MobPool mp{ s, env, N, s.m_next_bullet[env], s.m_num_bullets[env], s.m_next_expl[env], s.m_num_expl[env], ctx.leader() };
        int a_next = s.a_next_bullet[env], a_num = s.a_num_bullets[env];
        float a_timer = s.a_bullet_timer[env];
        bool alive = s.alive[env] != 0;
        const int nhaz = s.n_hazards[env];
        int hz_id[8];
        for (int k = 0; k < 8; k++) hz_id[k] = k < nhaz ? s.hazard_order[k * N + env] : 0;

        // barrier rectangles do not move within a step: fetched once, kept in registers (the loops over k are unrolled)
        float hz_x[8], hz_y[8];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const bool bar = k < nhaz && hz_id[k] >= 2;
            hz_x[k] = bar ? __fadd_rn(s.bar_x[(hz_id[k] - 2) * N + env], -0.1f) : 0.0f;
            hz_y[k] = bar ? __fadd_rn(s.bar_y[(hz_id[k] - 2) * N + env], -0.1f) : 0.0f;
        }
        auto hazard_rect_at = [&](int k, int id) {   // k-th hazard of the list (id = hz_id[k])
            if (id == 1) return Rect{ __fadd_rn(bx, -0.6f), __fadd_rn(by, -0.4f), 1.2f, 0.8f };
            return Rect{ hz_x[k], hz_y[k], 0.2f, 0.2f };
        };

        const float movement_x = (float)((action == 6 || action == 7 || action == 8) - (action == 0 || action == 1 || action == 2));
        const float movement_y = (float)((action == 2 || action == 5 || action == 8) - (action == 0 || action == 3 || action == 6));
        const bool fire = action == 9;
        bool agent_alive = true, boss_alive = true;

        for (int ss = 0; ss < SUB_STEPS; ss++) {
            // ================= System_Agent::update =================
            {
                pvx = __fadd_rn(pvx, __fmul_rn(__fmul_rn(0.5f, __fsub_rn(__fmul_rn(movement_x, 0.1f), pvx)), dt));
                pvy = __fadd_rn(pvy, __fmul_rn(__fmul_rn(0.5f, __fsub_rn(__fmul_rn(-movement_y, 0.1f), pvy)), dt));
                px = __fadd_rn(px, __fmul_rn(pvx, dt));
                py = __fadd_rn(py, __fmul_rn(pvy, dt));
                Rect wc{ __fadd_rn(px, -0.15f), __fadd_rn(py, -0.1f), 0.3f, 0.2f };
                const float sx1 = __fadd_rn(screen.x, screen.w), sy1 = __fadd_rn(screen.y, screen.h);
                if (wc.x < screen.x) { px = __fadd_rn(px, __fsub_rn(screen.x, wc.x)); pvx = 0.0f; }
                else if (__fadd_rn(wc.x, wc.w) > sx1) { px = __fadd_rn(px, __fsub_rn(sx1, __fadd_rn(wc.x, wc.w))); pvx = 0.0f; }
                if (wc.y < screen.y) { py = __fadd_rn(py, __fsub_rn(screen.y, wc.y)); pvy = 0.0f; }
                else if (__fadd_rn(wc.y, wc.h) > sy1) { py = __fadd_rn(py, __fsub_rn(sy1, __fadd_rn(wc.y, wc.h))); pvy = 0.0f; }
                wc = Rect{ __fadd_rn(px, -0.15f), __fadd_rn(py, -0.1f), 0.3f, 0.2f };

                if (fire) {
                    if (a_timer == 0.0f && a_num < AB) {
                        a_timer = 5.0f;
                        int i = a_next * N + env;
                        if (ctx.leader()) {
                            s.ab_vx[i] = 0.0f; s.ab_vy[i] = -0.1f;
                            s.ab_x[i] = px; s.ab_y[i] = py;
                            s.ab_frame[i] = 0.0f; s.ab_bouncing[i] = 0; s.ab_bounce_timer[i] = 0.0f;
                        }
                        a_next = (a_next + 1) % AB;
                        a_num++;
                    } else {
                        a_timer = fmaxf(0.0f, __fsub_rn(a_timer, dt));
                    }
                }
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    if (k >= nhaz) break;
                    if (check_collision(wc, hazard_rect_at(k, hz_id[k]))) { alive = false; break; }
                }

                // The agent's bullets, lane-parallel (ring slot i on lane i % nlanes). Pass 1: what would happen to each bullet
                // if it is visited (destroyed? bounces off the shield = one RNG draw? hits the unshielded boss?) — none of it
                // depends on the other bullets; the visited prefix follows from the destroyed bits; the bounce draws are made
                // in ring order by all lanes; pass 2 applies and stores.
                ctx.sync();   // the leader's store of a bullet fired above
                {
                    const int n0 = a_num;
                    auto slot = [&](int i) { return ((AB + a_next - 1 - i) % AB) * N + env; };
                    auto first_hazard = [&](const Rect& bw) {   // index into the hazard list of the first one hit, -1: none
#pragma unroll
                        for (int k = 0; k < 8; k++) {
                            if (k >= nhaz) break;
                            if (check_collision(bw, hazard_rect_at(k, hz_id[k]))) return k;
                        }
                        return -1;
                    };
                    uint64_t D = 0ull, R = 0ull, B = 0ull;
                    for (int i = ctx.lane; i < n0; i += ctx.nlanes) {
                        const int bi = slot(i);
                        float frame = s.ab_frame[bi];
                        if (frame == -1.0f) continue;
                        bool bouncing = s.ab_bouncing[bi] != 0;
                        float btimer = s.ab_bounce_timer[bi];
                        if (frame == 0.0f) {
                            const float x = s.ab_x[bi], y = s.ab_y[bi];
                            Rect bw{ __fsub_rn(x, 0.01f), __fsub_rn(y, 0.01f), 0.02f, 0.02f };
                            if (!check_collision(bw, screen)) frame = 5.0f;
                            else {
                                const int k = first_hazard(bw);
                                if (k >= 0) {
                                    if (hz_id[k] == 1) {
                                        if (phase_index % 2 == 0) { R |= 1ull << i; btimer = 10.0f; bouncing = true; }
                                        else { frame = 1.0f; B |= 1ull << i; }
                                    } else frame = 1.0f;
                                }
                            }
                        }
                        if (frame >= 5.0f || (bouncing && !(btimer > 0.0f))) D |= 1ull << i;
                    }
                    D = ctx.or64(D); R = ctx.or64(R); B = ctx.or64(B);
                    const int nvis = visited_prefix(D, n0);
                    const uint64_t vis = nvis >= 64 ? ~0ull : (1ull << nvis) - 1ull;
                    // bounce velocities: one draw per bouncing bullet, in ring order (uniform: every lane draws them all); the
                    // lane that owns the bullet stores the bounced vx right away and re-reads it in pass 2
                    for (uint64_t m = R & vis; m; m &= m - 1ull) {
                        const int i = __ffsll((long long)m) - 1;
                        const float r = rng.uniform_real(-1.0f, 1.0f);
                        if (i % ctx.nlanes == ctx.lane) s.ab_vx[slot(i)] = __fmul_rn(r, 0.05f);
                    }
                    {   // `if (hp > 0) hp--` per bullet that hits the unshielded boss
                        const int hits = __popcll(B & vis);
                        if (hits) hp = hp > hits ? hp - hits : 0;
                    }
                    for (int i = ctx.lane; i < nvis; i += ctx.nlanes) {
                        const int bi = slot(i);
                        float frame = s.ab_frame[bi];
                        if (frame == -1.0f) continue;
                        float x = s.ab_x[bi], y = s.ab_y[bi], vx = s.ab_vx[bi], vy = s.ab_vy[bi];
                        bool bouncing = s.ab_bouncing[bi] != 0;
                        float btimer = s.ab_bounce_timer[bi];
                        if (frame == 0.0f) {
                            Rect bw{ __fsub_rn(x, 0.01f), __fsub_rn(y, 0.01f), 0.02f, 0.02f };
                            if (!check_collision(bw, screen)) { vx = 0.0f; vy = 0.0f; frame = 5.0f; }
                            else {
                                const int k = first_hazard(bw);
                                if (k >= 0) {
                                    if (hz_id[k] == 1 && phase_index % 2 == 0) {   // shielded: bounce
                                        // (vx: the bounced value, stored by the draw loop above)
                                        vy = 0.05f;
                                        btimer = 10.0f; bouncing = true;
                                    } else { vx = 0.0f; vy = 0.0f; frame = 1.0f; }
                                }
                            }
                        }
                        x = __fadd_rn(x, __fmul_rn(vx, dt));
                        y = __fadd_rn(y, __fmul_rn(vy, dt));
                        bool destroy = false;
                        if (frame >= 5.0f) destroy = true;
                        else if (frame >= 1.0f) frame = __fadd_rn(frame, __fmul_rn(0.3f, dt));
                        if (bouncing) {
                            if (btimer > 0.0f) btimer = fmaxf(0.0f, __fsub_rn(btimer, dt));
                            else destroy = true;
                        }
                        if (destroy) frame = -1.0f;
                        s.ab_x[bi] = x; s.ab_y[bi] = y; s.ab_vx[bi] = vx; s.ab_vy[bi] = vy; s.ab_frame[bi] = frame;
                        s.ab_bouncing[bi] = bouncing; s.ab_bounce_timer[bi] = btimer;
                    }
                    a_num = n0 - __popcll(D & vis);
                    ctx.sync();
                }
                agent_alive = alive;
            }

            // ================= System_Mob_AI::update =================
            {
                boss_alive = true;
                const float shielded_phase_time = __fadd_rn(180.0f, __fmul_rn(rng.canonical(), SHIELD_JITTER));
                const Rect agent_rect{ __fadd_rn(px, -0.15f), __fadd_rn(py, -0.1f), 0.3f, 0.2f };
                if (phase_timer == 0.0f) {
                    weapon_index = rng.uniform_int(0, 3);
                    attack_timer = 0.0f;
                    hp = 3;
                }
                const float bullet_speed = BULLET_SPEED;
                auto fire_pattern = [&](int pattern) {
                    switch (pattern) {
                    case -1:
                        if (rng.canonical() < __fmul_rn(0.1f, dt)) {
                            float rot = (float)__dmul_rn(PI, (double)__fadd_rn(1.0f, rng.canonical()));
                            mp.fire(bx, by, rot, bullet_speed);
                        }
                        break;
                    case 0:
                        if (attack_timer >= 8.0f) {
                            attack_timer = 0.0f;
                            for (int i = 0; i < 5; i++) {
                                float rot = (float)__dadd_rn(__dmul_rn(PI, 1.5), __dmul_rn(__dmul_rn((double)(i - 2), PI), 0.125));
                                mp.fire(bx, by, rot, bullet_speed);
                            }
                        } else attack_timer = __fadd_rn(attack_timer, dt);
                        break;
                    case 1:
                        if (attack_timer >= 5.0f) {
                            attack_timer = 0.0f;
                            int k = 8;   // k = timer / 5 with timer just zeroed -> abs(8 - 0 % 16)
                            for (int i = 0; i < 4; i++) {
                                float f = __fadd_rn(1.25f, __fmul_rn((float)k, 0.0625f));
                                float rot = (float)__dadd_rn(__dmul_rn(PI, (double)f), __dmul_rn(__dmul_rn((double)i, PI), 0.5));
                                mp.fire(bx, by, rot, bullet_speed);
                            }
                        } else attack_timer = __fadd_rn(attack_timer, dt);
                        break;
                    case 2:
                        if (attack_timer >= 10.0f) {
                            attack_timer = 0.0f;
                            float offset = (float)__dmul_rn((double)__fmul_rn(rng.canonical(), 2.0f), PI);
                            for (int i = 0; i < 8; i++) {
                                float rot = (float)__dadd_rn(__dmul_rn(__dmul_rn(PI, 0.25), (double)i), (double)offset);
                                mp.fire(bx, by, rot, bullet_speed);
                            }
                        } else attack_timer = __fadd_rn(attack_timer, dt);
                        break;
                    default:
                        if (attack_timer >= 4.0f) {
                            attack_timer = 0.0f;
                            float rot = (float)__dmul_rn(PI, (double)__fadd_rn(1.0f, rng.canonical()));
                            mp.fire(bx, by, rot, bullet_speed);
                        } else attack_timer = __fadd_rn(attack_timer, dt);
                        break;
                    }
                };

                if (phase_index % 2 == 0) {
                    if (phase_timer >= shielded_phase_time) { phase_timer = 0.0f; phase_index++; }
                    else phase_timer = __fadd_rn(phase_timer, dt);
                    fire_pattern(weapon_index);
                } else {
                    if (phase_timer >= 300.0f) { phase_timer = 0.0f; phase_index++; }
                    else phase_timer = __fadd_rn(phase_timer, dt);
                    fire_pattern(-1);
                    if (hp == 0) {
                        if (expl_timer >= 8.0f) {   // show_damage
                            expl_timer = 0.0f;
                            float ex = __fadd_rn(rng.uniform_real(-0.5f, 0.5f), bx);
                            float ey = __fadd_rn(rng.uniform_real(-0.5f, 0.5f), by);
                            mp.explode(ex, ey);
                        } else expl_timer = __fadd_rn(expl_timer, dt);
                        if (damage_timer >= 80.0f) { damage_timer = 0.0f; phase_index++; hp = 3; }
                        else damage_timer = __fadd_rn(damage_timer, dt);
                    }
                }

                if (move_timer >= 70.0f) {
                    move_timer = 0.0f;
                    float r0 = rng.canonical();
                    float tx = __fmul_rn(__fmul_rn(__fmul_rn(__fsub_rn(__fmul_rn(r0, 2.0f), 1.0f), 0.5f), screen.w), 0.7f);
                    float r1 = rng.canonical();
                    float ty = __fmul_rn(__fmul_rn(__fsub_rn(__fmul_rn(__fsub_rn(__fmul_rn(r1, 2.0f), 1.0f), 0.5f), 0.3f), screen.h), 0.5f);
                    bvx = __fdiv_rn(__fsub_rn(tx, bx), 70.0f);
                    bvy = __fdiv_rn(__fsub_rn(ty, by), 70.0f);
                } else move_timer = __fadd_rn(move_timer, dt);
                bx = __fadd_rn(bx, __fmul_rn(bvx, dt));
                by = __fadd_rn(by, __fmul_rn(bvy, dt));

                // The boss's bullets, lane-parallel like the agent's: pass 1 finds, per bullet, whether a visit destroys it or
                // hits the agent (that visit ends the walk: `break`, common_systems.cpp:345-352), pass 2 applies and stores.
                ctx.sync();   // the leader's stores of the bullets fired above
                {
                    const int n0 = mp.num_bullets;
                    auto slot = [&](int i) { return ((MB + mp.next_bullet - 1 - i) % MB) * N + env; };
                    auto hits_barrier = [&](const Rect& bw) {
#pragma unroll
                        for (int k = 0; k < 8; k++) {
                            if (k >= nhaz) break;
                            if (hz_id[k] == 1) continue;
                            if (check_collision(bw, hazard_rect_at(k, hz_id[k]))) return true;
                        }
                        return false;
                    };
                    uint64_t D = 0ull, H = 0ull;
                    for (int i = ctx.lane; i < n0; i += ctx.nlanes) {
                        const int bi = slot(i);
                        float frame = s.mb_frame[bi];
                        if (frame == -1.0f) continue;
                        if (frame == 0.0f) {
                            const float x = s.mb_x[bi], y = s.mb_y[bi];
                            Rect bw{ __fsub_rn(x, 0.01f), __fsub_rn(y, 0.01f), 0.02f, 0.02f };
                            if (!check_collision(bw, screen)) frame = 5.0f;
                            else if (check_collision(bw, agent_rect)) { H |= 1ull << i; frame = 1.0f; }
                        }
                        if (frame >= 5.0f) D |= 1ull << i;
                    }
                    D = ctx.or64(D); H = ctx.or64(H);
                    int nvis = visited_prefix(D, n0);
                    uint64_t vis = nvis >= 64 ? ~0ull : (1ull << nvis) - 1ull;
                    int stop_at = -1;   // the visited bullet that hits the agent: stored unmoved, ends the walk
                    if (H & vis) { stop_at = __ffsll((long long)(H & vis)) - 1; nvis = stop_at + 1; vis = (1ull << nvis) - 1ull; alive = false; }
                    for (int i = ctx.lane; i < nvis; i += ctx.nlanes) {
                        const int bi = slot(i);
                        float frame = s.mb_frame[bi];
                        if (frame == -1.0f) continue;
                        float x = s.mb_x[bi], y = s.mb_y[bi], vx = s.mb_vx[bi], vy = s.mb_vy[bi];
                        if (frame == 0.0f) {
                            Rect bw{ __fsub_rn(x, 0.01f), __fsub_rn(y, 0.01f), 0.02f, 0.02f };
                            if (!check_collision(bw, screen)) { vx = 0.0f; vy = 0.0f; frame = 5.0f; }
                            else if (i == stop_at) { vx = 0.0f; vy = 0.0f; frame = 1.0f; }
                            else if (hits_barrier(bw)) { vx = 0.0f; vy = 0.0f; frame = 1.0f; }
                        }
                        if (i != stop_at) {
                            x = __fadd_rn(x, __fmul_rn(vx, dt));
                            y = __fadd_rn(y, __fmul_rn(vy, dt));
                            if (frame >= 5.0f) frame = -1.0f;
                            else if (frame >= 1.0f) frame = __fadd_rn(frame, __fmul_rn(0.3f, dt));
                        }
                        s.mb_x[bi] = x; s.mb_y[bi] = y; s.mb_vx[bi] = vx; s.mb_vy[bi] = vy; s.mb_frame[bi] = frame;
                    }
                    mp.num_bullets = n0 - __popcll(D & vis & ~(stop_at >= 0 ? 1ull << stop_at : 0ull));
                    ctx.sync();
                }
                for (int i = 0; i < mp.num_expl; i++) {
                    int ei = ((NEX + mp.next_expl - 1 - i) % NEX) * N + env;
                    float frame = s.ex_frame[ei];
                    if (frame == -1.0f) continue;
                    if (frame >= 4.0f) { mp.num_expl--; frame = -1.0f; }
                    else if (frame >= 0.0f) frame = __fadd_rn(frame, __fmul_rn(0.3f, dt));
                    if (ctx.leader()) s.ex_frame[ei] = frame;
                }
                if (phase_index >= 6) boss_alive = false;
            }
            if (!agent_alive || !boss_alive) break;
        }

        if (ctx.leader()) {
            s.px[env] = px; s.py[env] = py; s.pvx[env] = pvx; s.pvy[env] = pvy;
            s.bx[env] = bx; s.by[env] = by; s.bvx[env] = bvx; s.bvy[env] = bvy;
            s.phase_timer[env] = phase_timer; s.attack_timer[env] = attack_timer;
            s.phase_index[env] = phase_index; s.weapon_index[env] = weapon_index; s.hp[env] = hp;
            s.expl_timer[env] = expl_timer; s.damage_timer[env] = damage_timer; s.move_timer[env] = move_timer;
            s.m_next_bullet[env] = mp.next_bullet; s.m_num_bullets[env] = mp.num_bullets;
            s.m_next_expl[env] = mp.next_expl; s.m_num_expl[env] = mp.num_expl;
            s.a_next_bullet[env] = a_next; s.a_num_bullets[env] = a_num; s.a_bullet_timer[env] = a_timer;
            s.alive[env] = alive;
            c.mti[env] = rng.idx;
            c.sprites_valid[env] = 1;
        }
        // (!agent_alive) * -10.0f + (!boss_alive) * 10.0f
        *reward = __fadd_rn(__fmul_rn((float)(!agent_alive), -10.0f), __fmul_rn((float)(!boss_alive), 10.0f));
        return !agent_alive || !boss_alive;
    }

    // ---------------------------------------------------------------------------------------
    static PG2_DEV_NOINLINE void regenerate(const State& s, const CommonState& c, int env, WarpCtx& w) {
        const int N = s.N, lane = w.lane;
        const float half_w = 2.0f;   // camera_size / camera_scale * pixels_to_unit * 0.5f
        float player_x = __fmul_rn(__fmul_rn(__fdiv_rn(__fmul_rn(w.rng.uniform_real(-1.0f, 1.0f), 64.0f), 1.0f), PIXELS_TO_UNIT), 0.5f);
        int num_barriers = w.rng.uniform_int(1, 4);
        Rect* coll = w.alloc<Rect>(MAX_BAR);
        float* bxs = w.alloc<float>(MAX_BAR);
        float* bys = w.alloc<float>(MAX_BAR);
        uint8_t* btex = w.alloc<uint8_t>(MAX_BAR);
        int accepted = 0;
        for (int i = 0; i < num_barriers; i++) {
            float x = __fmul_rn(__fmul_rn(__fmul_rn(__fdiv_rn(__fmul_rn(w.rng.uniform_real(-1.0f, 1.0f), 64.0f), 1.0f), PIXELS_TO_UNIT), 0.5f), 0.9f);
            float y = __fsub_rn(half_w, w.rng.uniform_real(0.7f, 1.2f));
            Rect wc{ __fadd_rn(x, -0.1f), __fadd_rn(y, -0.1f), 0.2f, 0.2f };
            bool collided = false;
            for (int j = 0; j < i; j++)
                if (check_collision(wc, coll[j])) { collided = true; break; }
            __syncwarp();
            if (!collided) {
                int t = w.rng.uniform_int(0, 7);
                bxs[accepted] = x; bys[accepted] = y; btex[accepted] = (uint8_t)t;
                accepted++;
                coll[i] = wc;
            } else {
                coll[i] = Rect{ 0.0f, 0.0f, 0.0f, 0.0f };
            }
            __syncwarp();
        }
        int bg_index = w.rng.uniform_int(0, NUM_BG - 1);
        w.rng.canonical();   // current_background_offset_x (never read by render_game)
        w.rng.canonical();   // current_background_offset_y
        int a_ship = w.rng.uniform_int(0, 3);
        int a_bullet = w.rng.uniform_int(0, 2);
        int m_ship = w.rng.uniform_int(0, 3);
        int m_bullet = w.rng.uniform_int(0, 2);

        // iteration orders: hazard = {boss(1), barriers(2..)}, sprite_render = {barriers}
        USet<8, 64>* us = w.alloc<USet<8, 64>>(1);
        uint8_t* order = w.alloc<uint8_t>(8);
        __syncwarp();
        us->init(s.nb_hazard[env]);
        us->insert(1);
        for (int k = 0; k < accepted; k++) us->insert(2 + k);
        int nhaz = us->order(order);
        int nb_hazard = us->nb;
        __syncwarp();
        for (int k = lane; k < nhaz; k += WARP_LANES) s.hazard_order[k * N + env] = order[k];
        __syncwarp();
        us->init(s.nb_sprite[env]);
        for (int k = 0; k < accepted; k++) us->insert(2 + k);
        int nspr = us->order(order);
        int nb_sprite = us->nb;
        __syncwarp();
        for (int k = lane; k < nspr; k += WARP_LANES) s.sprite_order[k * N + env] = order[k];

        for (int k = lane; k < accepted; k += WARP_LANES) {
            s.bar_x[k * N + env] = bxs[k]; s.bar_y[k * N + env] = bys[k]; s.bar_tex[k * N + env] = btex[k];
        }
        for (int k = lane; k < MB; k += WARP_LANES) s.mb_frame[k * N + env] = -1.0f;
        for (int k = lane; k < AB; k += WARP_LANES) s.ab_frame[k * N + env] = -1.0f;
        for (int k = lane; k < NEX; k += WARP_LANES) s.ex_frame[k * N + env] = -1.0f;
        if (lane == 0) {
            s.px[env] = player_x; s.py[env] = half_w; s.pvx[env] = 0.0f; s.pvy[env] = 0.0f;
            s.bx[env] = 0.0f; s.by[env] = 0.0f; s.bvx[env] = 0.0f; s.bvy[env] = 0.0f;
            s.phase_timer[env] = 0.0f; s.phase_index[env] = 0; s.weapon_index[env] = 0; s.attack_timer[env] = 0.0f; s.hp[env] = 0;
            s.m_next_bullet[env] = 0; s.m_next_expl[env] = 0; s.m_num_bullets[env] = 0; s.m_num_expl[env] = 0;
            s.expl_timer[env] = 0.0f; s.damage_timer[env] = 0.0f; s.move_timer[env] = 0.0f;
            s.m_ship[env] = m_ship; s.m_bullet_tex[env] = m_bullet;
            s.a_next_bullet[env] = 0; s.a_num_bullets[env] = 0; s.a_bullet_timer[env] = 0.0f;
            s.a_ship[env] = a_ship; s.a_bullet_tex[env] = a_bullet; s.alive[env] = 1;
            s.num_barriers[env] = accepted; s.n_hazards[env] = nhaz;
            s.nb_hazard[env] = nb_hazard; s.nb_sprite[env] = nb_sprite;
            s.bg_index[env] = bg_index;
            c.sprites_valid[env] = 0;
        }
    }

    // ---------------------------------------------------------------------------------------
    static PG2_DEV int tile_class(uint32_t) { return 0; }

    // NOTE: bullets / explosions that died keep frame == -1 and are skipped exactly like the reference does;
    // the pools are NOT cleared by reset() in the reference either (only the counters are), but a slot is
    // only ever visited while it lies within `num` positions behind `next`, i.e. after being rewritten.
    template <class F>
    static PG2_DEV_NOINLINE void build_frame(const State& s, const CommonState& c, int env, F& f, const TexInfo* tex) {
        const int N = s.N;
        const Camera cam{ 0.0f, 0.0f, __fdiv_rn(__fmul_rn(1.0f, f.view_w), 64.0f), f.view_w, f.view_h };   // game_zoom * width / obs_width
        const double PI = 3.14159265358979323846;
        {   // background only, no tile layer (bossfight.cpp:424: centred, scaled to the camera height)
            const int bg = T_BG0 + s.bg_index[env];
            const float sc = __fdiv_rn(__fmul_rn(__fdiv_rn(1.0f, (float)tex[bg].h), cam.h), cam.scale);
            const float half_x = __fmul_rn(__fdiv_rn(-cam.w, cam.scale), 0.5f), half_y = __fmul_rn(__fdiv_rn(-cam.h, cam.scale), 0.5f);
            build_tile_layer(f, cam, tex, 1, 0, 0, 0, 0, [](int) { return 0; }, [](int, int) { return (int)NO_TILE; }, bg, half_x, half_y, sc);
        }
        const int m_num = s.m_num_bullets[env], m_next = s.m_next_bullet[env];
        const int e_num = s.m_num_expl[env], e_next = s.m_next_expl[env];
        const int a_num = s.a_num_bullets[env], a_next = s.a_next_bullet[env];
        const int nspr = c.sprites_valid[env] ? s.num_barriers[env] : 0;
        const bool shield = s.phase_index[env] % 2 == 0;
        const float bx = s.bx[env], by = s.by[env];
        // submission order: boss bullets, boss, shield, explosions, barrier sprites, agent bullets, ship
        const int o_boss = m_num, o_shield = o_boss + 1, o_expl = o_shield + 1, o_spr = o_expl + e_num,
                  o_ab = o_spr + nspr, o_ship = o_ab + a_num;
        emit_post_blits(f, tex, o_ship + 1, [&](int k, BlitReq& b, BlitRot& rot) {
            if (k < o_boss) {
                int bi = ((MB + m_next - 1 - k) % MB) * N + env;
                float frame = s.mb_frame[bi];
                if (frame == -1.0f) return;
                int t = frame == 0.0f ? T_BULLET0 + s.m_bullet_tex[env] : T_EXPL0 + f2i(__fsub_rn(frame, 1.0f));
                const float size = 0.1f;
                float x = __fsub_rn(__fmul_rn(s.mb_x[bi], UNIT_TO_PIXELS), __fmul_rn(__fmul_rn(size, (float)tex[t].w), 0.5f));
                float y = __fsub_rn(__fmul_rn(s.mb_y[bi], UNIT_TO_PIXELS), __fmul_rn(__fmul_rn(size, (float)tex[t].h), 0.5f));
                float rotation = (float)__dadd_rn((double)s.mb_rot[bi], __dmul_rn(PI, 0.5));
                b.rotated(t, x, y, cam, rotation, size, 1.0f, &rot);
            } else if (k == o_boss || k == o_shield) {
                if (k == o_shield && !shield) return;
                int t = k == o_boss ? T_BOSS0 + s.m_ship[env] : T_SHIELD;
                const float size = 0.25f;
                float x = __fsub_rn(__fmul_rn(bx, UNIT_TO_PIXELS), __fmul_rn(__fmul_rn(size, (float)tex[t].w), 0.5f));
                float y = __fsub_rn(__fmul_rn(by, UNIT_TO_PIXELS), __fmul_rn(__fmul_rn(size, (float)tex[t].h), 0.5f));
                b.plain(t, x, y, cam, size, k == o_boss ? 1.0f : 0.7f);
            } else if (k < o_spr) {
                int ei = ((NEX + e_next - 1 - (k - o_expl)) % NEX) * N + env;
                float frame = s.ex_frame[ei];
                if (frame == -1.0f) return;
                int t = T_EXPL0 + f2i(frame);
                const float size = 0.3f;
                float x = __fsub_rn(__fmul_rn(s.ex_x[ei], UNIT_TO_PIXELS), __fmul_rn(__fmul_rn(size, (float)tex[t].w), 0.5f));
                float y = __fsub_rn(__fmul_rn(s.ex_y[ei], UNIT_TO_PIXELS), __fmul_rn(__fmul_rn(size, (float)tex[t].h), 0.5f));
                b.plain(t, x, y, cam, size);
            } else if (k < o_ab) {
                int id = s.sprite_order[sort_perm(nspr, k - o_spr) * N + env] - 2;
                int t = T_BARRIER0 + s.bar_tex[id * N + env];
                float x = __fmul_rn(__fadd_rn(s.bar_x[id * N + env], -0.15f), UNIT_TO_PIXELS);
                float y = __fmul_rn(__fadd_rn(s.bar_y[id * N + env], -0.15f), UNIT_TO_PIXELS);
                float sc = __fdiv_rn(__fmul_rn(__fmul_rn(1.0f, 0.3f), UNIT_TO_PIXELS), (float)tex[t].w);
                b.plain(t, x, y, cam, sc);
            } else if (k < o_ship) {
                int bi = ((AB + a_next - 1 - (k - o_ab)) % AB) * N + env;
                float frame = s.ab_frame[bi];
                if (frame == -1.0f) return;
                int t = frame == 0.0f ? T_BULLET0 + s.a_bullet_tex[env] : T_EXPL0 + f2i(__fsub_rn(frame, 1.0f));
                const float size = 0.05f;
                float x = __fsub_rn(__fmul_rn(s.ab_x[bi], UNIT_TO_PIXELS), __fmul_rn(__fmul_rn(size, (float)tex[t].w), 0.5f));
                float y = __fsub_rn(__fmul_rn(s.ab_y[bi], UNIT_TO_PIXELS), __fmul_rn(__fmul_rn(size, (float)tex[t].h), 0.5f));
                b.plain(t, x, y, cam, size);
            } else {
                int t = T_PLAYER0 + s.a_ship[env];
                const float size = 0.05f;
                float x = __fsub_rn(__fmul_rn(s.px[env], UNIT_TO_PIXELS), __fmul_rn(__fmul_rn(size, (float)tex[t].w), 0.5f));
                float y = __fsub_rn(__fmul_rn(s.py[env], UNIT_TO_PIXELS), __fmul_rn(__fmul_rn(size, (float)tex[t].h), 0.5f));
                b.plain(t, x, y, cam, size);
            }
        });
    }
};
using BossFight = BossFightT<1>;

}  // namespace pg2
