// oracle/ref_shim.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Thin C-ABI shim around the UNMODIFIED upstream AVIR / LANCIR headers, which are
// compiled from where they lie (the directory passed as -I by oracle/Makefile,
// normally /root/reference).  No upstream source is copied into this repository.
// The resulting oracle/_ref/libavir_ref.so is
//   * the parity oracle for tests/ (whole-image outputs + plan dumps),
//   * the "reference" CPU baseline that bench.py times on the host cores.
//
// Pinned build flags: g++ -O2 -mavx2 -ffp-contract=off.
//
// "Spy" filter-step classes below derive from the upstream filter-step classes and
// forward every call to them; they only record the plan (step list, taps,
// resize positions) that upstream resizeImage() built, on the first scanline of the
// H pass and of the V pass.  The arithmetic is entirely upstream's.

#include <cstdint>
#include <cstring>
#include <cstdio>
#include <vector>
#include <thread>
#include <mutex>
#include <condition_variable>
#include <atomic>
#include <set>

#include "avir.h"
#include "avir_float4_sse.h"
#include "avir_float8_avx.h"
#include "lancir.h"

namespace {

// ---------------------------------------------------------------- plan recorder

struct RecStep
{
	int kind; // 0 = FIR (doFilter), 1 = upsample, 2 = resize, 3 = resize2
	int ResampleFactor;
	int FltLatency;
	int EdgePixelCount;
	int InLen, InPrefix, InSuffix, OutLen, OutPrefix, OutSuffix;
	int FltOrigLen; // >0 => filterless upsample
	int BankFilterLen, BankOrder, BankFracCount;
	std::vector< float > Flt;
	std::vector< int > SrcPosInt, fti, fl;
	std::vector< float > x;
	std::vector< int > used_fti; // sorted distinct phases
	std::vector< float > bank; // used phases, FilterLen*(Order+1) floats each
};

struct Recorder
{
	bool armed = false;
	int pass = -1; // 0 = H, 1 = V
	int seen[ 2 ] = { 0, 0 }; // scanlines seen per pass
	std::vector< RecStep > steps[ 2 ];
};

thread_local Recorder* g_rec = nullptr;

inline float atom0( const float& v ) { return v; }
inline float atom0( const avir::float4& v ) { float t[ 4 ]; v.storeu( t ); return t[ 0 ]; }

template< class Step >
void recordStep( const Step& s, const int kind )
{
	Recorder* const r = g_rec;

	if( r == nullptr || !r -> armed || r -> pass < 0 ||
		r -> seen[ r -> pass ] != 1 )
	{
		return;
	}

	RecStep rs;
	rs.kind = kind;
	rs.ResampleFactor = s.ResampleFactor;
	rs.FltLatency = ( kind >= 2 ? 0 : s.FltLatency );
	rs.EdgePixelCount = ( kind == 0 ? s.EdgePixelCount : 0 );
	rs.InLen = s.InLen;
	rs.InPrefix = s.InPrefix;
	rs.InSuffix = s.InSuffix;
	rs.OutLen = s.OutLen;
	rs.OutPrefix = ( kind == 1 ? s.OutPrefix : 0 );
	rs.OutSuffix = ( kind == 1 ? s.OutSuffix : 0 );
	rs.FltOrigLen = s.FltOrig.getCapacity();
	rs.BankFilterLen = rs.BankOrder = rs.BankFracCount = 0;

	if( kind < 2 )
	{
		for( int i = 0; i < s.Flt.getCapacity(); i++ )
		{
			rs.Flt.push_back( atom0( s.Flt[ i ]));
		}
	}
	else
	{
		const int FL = s.FltBank -> getFilterLen();
		const int ord = s.FltBank -> getOrder();
		rs.BankFilterLen = FL;
		rs.BankOrder = ord;
		rs.BankFracCount = s.FltBank -> getFracCount();
		std::set< int > used;

		for( int i = 0; i < s.OutLen; i++ )
		{
			const auto& rp = (*s.RPosBuf)[ i ];
			rs.SrcPosInt.push_back( rp.SrcPosInt );
			rs.fti.push_back( rp.fti );
			rs.x.push_back( (float) rp.x );
			rs.fl.push_back( kind == 3 ? rp.fl : FL );
			used.insert( rp.fti );
		}

		for( int f : used )
		{
			rs.used_fti.push_back( f );
			const auto* p = s.FltBank -> getFilterConst( f );

			for( int i = 0; i < FL * ( ord + 1 ); i++ )
			{
				rs.bank.push_back( atom0( p[ i ]));
			}
		}
	}

	r -> steps[ r -> pass ].push_back( rs );
}

template< class Base >
class SpyStep : public Base
{
public:
	template< class Tin, class fpt >
	void packScanline( const Tin* ip, fpt* const op, const int l ) const
	{
		if( g_rec != nullptr && g_rec -> armed )
		{
			g_rec -> pass = 0;
			g_rec -> seen[ 0 ]++;
		}

		Base :: packScanline( ip, op, l );
	}

	template< class fpt >
	void convertVtoH( const fpt* ip, fpt* op, const int SrcLen,
		const int SrcIncr ) const
	{
		if( g_rec != nullptr && g_rec -> armed )
		{
			g_rec -> pass = 1;
			g_rec -> seen[ 1 ]++;
		}

		Base :: convertVtoH( ip, op, SrcLen, SrcIncr );
	}

	template< class A, class B >
	void doUpsample( A&& Src, B&& Dst ) const
	{
		recordStep( *this, 1 );
		Base :: doUpsample( Src, Dst );
	}

	template< class A, class B >
	void doFilter( A&& Src, B&& Dst, const int DstIncr ) const
	{
		recordStep( *this, 0 );
		Base :: doFilter( Src, Dst, DstIncr );
	}

	template< class A, class B, class D >
	void doResize( A&& SrcLine, B&& DstLine, const int DstLineIncr,
		D&& xx ) const
	{
		recordStep( *this, 2 );
		Base :: doResize( SrcLine, DstLine, DstLineIncr, xx );
	}

	template< class A, class B, class D >
	void doResize2( A&& SrcLine, B&& DstLine, const int DstLineIncr,
		D&& xx ) const
	{
		recordStep( *this, this -> Vars -> IsResize2 ? 3 : 2 );
		Base :: doResize2( SrcLine, DstLine, DstLineIncr, xx );
	}
};

template< class fpbase >
class fpclass_spy : public fpbase
{
public:
	typedef SpyStep< typename fpbase :: CFilterStep > CFilterStep;
};

typedef fpclass_spy< avir::fpclass_def< float > > spy_def;
typedef fpclass_spy< avir::fpclass_float4 > spy_float4;
typedef fpclass_spy< avir::fpclass_float8_dil > spy_float8_dil;

// ---------------------------------------------------------------- thread pool

// std::thread pool for the CPU baseline.  Upstream deals scanlines round-robin to
// `ThreadCount` workloads and runs workload 0 on the calling thread
// (avir.h:4861-4894); workloads 1..N-1 run here.
class StdThreadPool : public avir::CImageResizerThreadPool
{
public:
	explicit StdThreadPool( const int n )
		: Count( n < 1 ? 1 : n )
		, Gen( 0 )
		, Pending( 0 )
		, Quit( false )
	{
		for( int i = 1; i < Count; i++ )
		{
			Threads.emplace_back( [ this, i ]() { run( i - 1 ); } );
		}
	}

	~StdThreadPool() override
	{
		{
			std::lock_guard< std::mutex > lk( Mx );
			Quit = true;
			Gen++;
		}

		Cv.notify_all();

		for( auto& t : Threads )
		{
			t.join();
		}
	}

	int getSuggestedWorkloadCount() const override { return Count; }

	void addWorkload( CWorkload* const w ) override { Work.push_back( w ); }

	void startAllWorkloads() override
	{
		{
			std::lock_guard< std::mutex > lk( Mx );
			Pending = (int) Work.size();
			Gen++;
		}

		Cv.notify_all();
	}

	void waitAllWorkloadsToFinish() override
	{
		std::unique_lock< std::mutex > lk( Mx );
		CvDone.wait( lk, [ this ]() { return Pending == 0; } );
	}

	void removeAllWorkloads() override { Work.clear(); }

private:
	int Count;
	std::vector< std::thread > Threads;
	std::vector< CWorkload* > Work;
	std::mutex Mx;
	std::condition_variable Cv, CvDone;
	long Gen;
	int Pending;
	bool Quit;

	void run( const int slot )
	{
		long seen = 0;

		while( true )
		{
			CWorkload* w = nullptr;
			{
				std::unique_lock< std::mutex > lk( Mx );
				Cv.wait( lk, [ & ]() { return Gen != seen; } );
				seen = Gen;

				if( Quit )
				{
					return;
				}

				if( slot < (int) Work.size() )
				{
					w = Work[ slot ];
				}
			}

			if( w != nullptr )
			{
				w -> process();
				std::lock_guard< std::mutex > lk( Mx );

				if( --Pending == 0 )
				{
					CvDone.notify_all();
				}
			}
		}
	}
};

// ---------------------------------------------------------------- dispatch

struct Call
{
	const void* src; int sw, sh, sls;
	void* dst; int nw, nh, C;
	double k; int resbits, srcbits;
	double ox, oy; int gamma, alpha, buildmode, nthreads;
	int params; // 0 Def, 1 ULR, 2 LR, 3 Low, 4 High, 5 Ultra
};

avir::CImageResizerParams mkParams( const int id )
{
	switch( id )
	{
		case 1: return avir::CImageResizerParamsULR();
		case 2: return avir::CImageResizerParamsLR();
		case 3: return avir::CImageResizerParamsLow();
		case 4: return avir::CImageResizerParamsHigh();
		case 5: return avir::CImageResizerParamsUltra();
		default: return avir::CImageResizerParamsDef();
	}
}

template< class fpclass, class Tin, class Tout >
void run3( const Call& c )
{
	avir::CImageResizer< fpclass > rs( c.resbits, c.srcbits, mkParams( c.params ));
	avir::CImageResizerVars v;
	v.ox = c.ox;
	v.oy = c.oy;
	v.UseSRGBGamma = ( c.gamma != 0 );
	v.AlphaIndex = c.alpha;
	v.BuildMode = c.buildmode;
	StdThreadPool* tp = nullptr;

	if( c.nthreads > 1 )
	{
		tp = new StdThreadPool( c.nthreads );
		v.ThreadPool = tp;
	}

	rs.resizeImage( (const Tin*) c.src, c.sw, c.sh, c.sls, (Tout*) c.dst,
		c.nw, c.nh, c.C, c.k, &v );

	delete tp;
}

template< class fpclass, class Tin >
int run2( const Call& c, const int tout )
{
	switch( tout )
	{
		case 0: run3< fpclass, Tin, uint8_t >( c ); return 0;
		case 1: run3< fpclass, Tin, uint16_t >( c ); return 0;
		case 2: run3< fpclass, Tin, float >( c ); return 0;
		case 3: run3< fpclass, Tin, double >( c ); return 0;
	}

	return -1;
}

template< class fpclass >
int run1( const Call& c, const int tin, const int tout )
{
	switch( tin )
	{
		case 0: return run2< fpclass, uint8_t >( c, tout );
		case 1: return run2< fpclass, uint16_t >( c, tout );
		case 2: return run2< fpclass, float >( c, tout );
		case 3: return run2< fpclass, double >( c, tout );
	}

	return -1;
}

int run0( const int fpc, const Call& c, const int tin, const int tout,
	const bool spy )
{
	if( spy )
	{
		switch( fpc )
		{
			case 0: return run1< spy_def >( c, tin, tout );
			case 1: return run1< spy_float4 >( c, tin, tout );
			case 2: return run1< spy_float8_dil >( c, tin, tout );
		}
	}
	else
	{
		switch( fpc )
		{
			case 0: return run1< avir::fpclass_def< float > >( c, tin, tout );
			case 1: return run1< avir::fpclass_float4 >( c, tin, tout );
			case 2: return run1< avir::fpclass_float8_dil >( c, tin, tout );
			// the same classes composed with upstream's error-diffusion ditherer
			case 3: return run1< avir::fpclass_def< float, float,
				avir::CImageResizerDithererErrdINL< float > > >( c, tin, tout );
			case 4: return run1< avir::fpclass_def< avir::float4, float,
				avir::CImageResizerDithererErrdINL< avir::float4 > > >( c, tin, tout );
			case 5: return run1< avir::fpclass_def_dil< float, avir::float8,
				avir::CImageResizerDithererErrdDIL< float, avir::float8 > > >( c, tin, tout );
		}
	}

	return -1;
}

void put( std::vector< double >& o, double v ) { o.push_back( v ); }

} // namespace

extern "C" {

// fpclass: 0 = fpclass_def<float>, 1 = fpclass_float4, 2 = fpclass_float8_dil.
// tin/tout: 0 = uint8_t, 1 = uint16_t, 2 = float, 3 = double.
int avir_ref_resize( int fpclass, int tin, int tout,
	const void* src, int sw, int sh, int sls, void* dst, int nw, int nh, int C,
	double k, int resbits, int srcbits, double ox, double oy, int gamma,
	int alpha, int buildmode, int nthreads, int params )
{
	const Call c = { src, sw, sh, sls, dst, nw, nh, C, k, resbits, srcbits,
		ox, oy, gamma, alpha, buildmode, nthreads, params };

	return run0( fpclass, c, tin, tout, false );
}

// Runs upstream resizeImage() through the spy fpclass and serialises the plan it
// built into `out` (doubles; every float is exactly representable).  Returns the
// number of doubles required (call with cap = 0 to size).  Layout:
//   for pass in (H, V): nsteps, then per step:
//     kind RF lat edge InLen InPrefix InSuffix OutLen OutPrefix OutSuffix FltOrigLen
//     nFlt Flt[...]
//     bankFL bankOrder bankFracCount nPos (SrcPosInt fti x fl)[nPos]
//     nUsed (fti taps[FL*(ord+1)])[nUsed]
long avir_ref_plan( int fpclass, int tin, int tout,
	const void* src, int sw, int sh, int sls, void* dst, int nw, int nh, int C,
	double k, int resbits, int srcbits, double ox, double oy, int gamma,
	int alpha, int buildmode, int params, double* out, long cap )
{
	Recorder rec;
	rec.armed = true;
	g_rec = &rec;

	const Call c = { src, sw, sh, sls, dst, nw, nh, C, k, resbits, srcbits,
		ox, oy, gamma, alpha, buildmode, 1, params };

	const int r = run0( fpclass, c, tin, tout, true );
	g_rec = nullptr;

	if( r != 0 )
	{
		return -1;
	}

	std::vector< double > o;

	for( int p = 0; p < 2; p++ )
	{
		put( o, (double) rec.steps[ p ].size() );

		for( const RecStep& s : rec.steps[ p ])
		{
			put( o, s.kind ); put( o, s.ResampleFactor ); put( o, s.FltLatency );
			put( o, s.EdgePixelCount ); put( o, s.InLen ); put( o, s.InPrefix );
			put( o, s.InSuffix ); put( o, s.OutLen ); put( o, s.OutPrefix );
			put( o, s.OutSuffix ); put( o, s.FltOrigLen );
			put( o, (double) s.Flt.size() );

			for( float f : s.Flt ) put( o, f );

			put( o, s.BankFilterLen ); put( o, s.BankOrder );
			put( o, s.BankFracCount ); put( o, (double) s.fti.size() );

			for( size_t i = 0; i < s.fti.size(); i++ )
			{
				put( o, s.SrcPosInt[ i ]); put( o, s.fti[ i ]);
				put( o, s.x[ i ]); put( o, s.fl[ i ]);
			}

			put( o, (double) s.used_fti.size() );
			const size_t fs = (size_t) s.BankFilterLen * ( s.BankOrder + 1 );

			for( size_t u = 0; u < s.used_fti.size(); u++ )
			{
				put( o, s.used_fti[ u ]);

				for( size_t i = 0; i < fs; i++ ) put( o, s.bank[ u * fs + i ]);
			}
		}
	}

	if( (long) o.size() <= cap && out != nullptr )
	{
		memcpy( out, o.data(), o.size() * sizeof( double ));
	}

	return (long) o.size();
}

// CLancIR::resizeImage (lancir.h:386).  Returns upstream's return value.
int lancir_ref_resize( int tin, int tout, const void* src, int sw, int sh,
	void* dst, int nw, int nh, int C, int srcssize, int newssize,
	double kx, double ky, double ox, double oy, double la )
{
	avir::CLancIR r;
	avir::CLancIRParams p( srcssize, newssize, kx, ky, ox, oy );
	p.la = la;

#define LR( TI, TO ) return r.resizeImage( (const TI*) src, sw, sh, (TO*) dst, nw, nh, C, &p )

	switch( tin * 3 + tout )
	{
		case 0: LR( uint8_t, uint8_t );
		case 1: LR( uint8_t, uint16_t );
		case 2: LR( uint8_t, float );
		case 3: LR( uint16_t, uint8_t );
		case 4: LR( uint16_t, uint16_t );
		case 5: LR( uint16_t, float );
		case 6: LR( float, uint8_t );
		case 7: LR( float, uint16_t );
		case 8: LR( float, float );
	}

#undef LR

	return -1;
}

// Runs `reps` LANCIR resizes on `nthreads` independent CLancIR objects (one per
// thread, upstream contract lancir.h:319-324), each thread on its own dst copy.
int lancir_ref_resize_mt( int tin, int tout, const void* src, int sw, int sh,
	void* dst, size_t dst_stride_bytes, int nw, int nh, int C, int nthreads )
{
	std::vector< std::thread > th;
	std::atomic< int > bad( 0 );

	for( int t = 0; t < nthreads; t++ )
	{
		th.emplace_back( [ =, &bad ]()
		{
			void* d = (char*) dst + (size_t) t * dst_stride_bytes;

			if( lancir_ref_resize( tin, tout, src, sw, sh, d, nw, nh, C, 0, 0,
				0.0, 0.0, 0.0, 0.0, 3.0 ) != nh )
			{
				bad++;
			}
		} );
	}

	for( auto& t : th ) t.join();

	return bad.load();
}

const char* avir_ref_version()
{
	return "avir " AVIR_VERSION " + lancir.h"
		" (g++ -O2 -mavx2 -ffp-contract=off)";
}

} // extern "C"
