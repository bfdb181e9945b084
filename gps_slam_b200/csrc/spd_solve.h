// Small symmetric-positive-definite solves of the tracker's normal equations (3x3 when only rotation or translation is
// estimated, 6x6 otherwise) and the determinant its pose-quality score needs.  Square-root-free Cholesky (A = L D L^T) with
// the dimension as a template parameter, kept as three separate pieces -- the un-normalised column entries g(r,c), the unit
// lower factor l(r,c) = g(r,c) / d(c), and the pivots d(c) -- instead of one overwritten matrix.
//
// The tracker's result has to reproduce the reference's LM trajectory (behaviour studied in InfiniTAM/ORUtils/Cholesky.h:16-86:
// which entries are subtracted in which order, one reciprocal per pivot, divide-by-pivot between the two substitutions), so the
// rounding order is part of the contract and is stated per step:
//   g(r,c) = A[c + r N] - g(c,0) l(r,0) - g(c,1) l(r,1) - ...        (left to right, k < c)
//   d(c)   = g(c,c),  l(r,c) = g(r,c) * (1 / d(c))
//   forward  y(i) = b(i) - l(i,0) y(0) - ... ;  y(i) /= d(i) ;  backward  x(i) = y(i) - l(i+1,i) x(i+1) - ... (ascending index)
//   det(A)^2 = (d(0) d(1) ... )^2
#pragma once

#ifdef __CUDACC__
#define SPD_HD __host__ __device__
#else
#define SPD_HD
#endif

template <int N>
struct SpdFactor
{
    float g[N][N], l[N][N], d[N];

    SPD_HD explicit SpdFactor(const float *A)
    {
        for (int c = 0; c < N; c++)
        {
            float rpivot = 1.0f;
            for (int r = c; r < N; r++)
            {
                float acc = A[c + r * N];
                for (int k = 0; k < c; k++)
                    acc -= g[c][k] * l[r][k];
                if (r == c)
                {
                    d[c] = acc;
                    rpivot = 1.0f / acc;
                }
                else
                {
                    g[r][c] = acc;
                    l[r][c] = acc * rpivot;
                }
            }
        }
    }

    SPD_HD void solve(const float *b, float *x) const
    {
        float y[N];
        for (int i = 0; i < N; i++)
        {
            float acc = b[i];
            for (int k = 0; k < i; k++)
                acc -= l[i][k] * y[k];
            y[i] = acc;
        }
        for (int i = 0; i < N; i++)
            y[i] /= d[i];
        for (int i = N - 1; i >= 0; i--)
        {
            float acc = y[i];
            for (int k = i + 1; k < N; k++)
                acc -= l[k][i] * x[k];
            x[i] = acc;
        }
    }

    SPD_HD float det_squared() const
    {
        float prod = d[0];
        for (int i = 1; i < N; i++)
            prod *= d[i];
        return prod * prod;
    }
};
